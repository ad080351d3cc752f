"""TEST INFRASTRUCTURE -- float64 numpy oracle for FK / sphere-model distances / masks / do_dq.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product package (cbf-inc-rrt_b200/) never does.

Parity status: **parity unpinned** against pybullet 3.2.5 (not installed, no network, reference ships no
golden vectors -- SURVEY.md section 8c).  What IS pinned: the chain constants come straight from the URDF
text, and FK at q=0 / q0 matches the hand-computed values listed in SURVEY.md section 8c
(tests/test_oracle_fk.py).

This is an independent derivation (4x4 homogeneous transforms built from the raw URDF xyz/rpy in
float64, un-snapped) used to bound the error of the float32 "deterministic contract" C oracle
(oracle/cbfrrt_oracle.c) and, through it, of the CUDA kernels.

Reference semantics restated here
  FK                         environment/basic_robot.py:55-68  (URDF link frames, base at origin)
  safe (soft) mask           neural_cbf/systems/arm_dynamics.py:169-182,189-210; basic_robot.py:89-98
  unsafe (hard) mask         neural_cbf/systems/arm_dynamics.py:184-187,212-232; environment/arm_env.py:82-84
  observation o, do/dq       neural_cbf/systems/arm_mindis.py:56-90
"""
import json
import os

import numpy as np

_SPEC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spec", "robot_models.json")


def load_models():
    with open(_SPEC) as f:
        return json.load(f)


def _rpy(r, p, y):
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def fk(model, q):
    """q: (n,) -> list over pybullet links of (pos(3), R(3,3)); basic_robot.py:55-62 semantics."""
    q = np.asarray(q, dtype=np.float64)
    frames = {-1: np.eye(4)}
    for l in model["links"]:
        T = np.eye(4)
        T[:3, :3] = _rpy(*l["rpy64"])
        T[:3, 3] = l["t64"]
        if l["qidx"] >= 0:
            a = q[l["qidx"]]
            Rz = np.eye(4)
            Rz[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
            T = T @ Rz
        frames[l["index"]] = frames[l["parent"]] @ T
    return frames


def sphere_centres(model, frames):
    C = np.zeros((len(model["spheres"]), 3))
    for i, s in enumerate(model["spheres"]):
        T = frames[s["link"]]
        C[i] = T[:3, :3] @ np.array(s["c"], dtype=np.float64) + T[:3, 3]
    return C


def sd_box(p, c, h):
    """signed distance of point p to an axis-aligned box (centre c, half extents h) and its gradient."""
    a = np.abs(p - c) - h
    out = np.maximum(a, 0.0)
    dout = np.linalg.norm(out)
    inside = min(max(a[0], a[1], a[2]), 0.0)
    sd = dout + inside
    sgn = np.where(p - c >= 0, 1.0, -1.0)
    if dout > 0:
        n = out / dout * sgn
    else:
        k = int(np.argmax(a))
        n = np.zeros(3)
        n[k] = sgn[k]
    return sd, n


def evaluate(model, q, boxes, spheres_obs, floor, thr_soft, thr_hard=0.0):
    """Full per-state geometric evaluation in float64.

    boxes: (Ob,6) [cx,cy,cz,hx,hy,hz]; spheres_obs: (Os,4) [cx,cy,cz,r]; floor: bool.
    Obstacle index convention = reference's env.obstacle_ids order: floor first (if present), then boxes,
    then (extension) sphere obstacles.
    Returns dict(safe, unsafe, o, arg_sphere, arg_obs, do_dq, self_min, soft_min, hard_min).
    """
    frames = fk(model, q)
    C = sphere_centres(model, frames)
    n = model["n"]
    off = 1 if floor else 0
    best = dict(o=np.inf, arg_s=-1, arg_o=-1, nrm=None)
    soft_min = np.inf
    hard_min = np.inf
    first_of_link = {}
    for i, s in enumerate(model["spheres"]):
        first_of_link.setdefault(s["link"], i)
    for i, s in enumerate(model["spheres"]):
        link = s["link"]
        r = s["r"]
        cands = []
        if floor:
            fc = model["floor_const"]
            if link == fc["link"]:
                # link 0 only spins about world z: its floor distance is the exact kinematic constant of the
                # spec (tools/gen_robot_model.py), attached to the link's first sphere; the other spheres of
                # that link skip the floor.
                if i == first_of_link[link]:
                    cands.append((0, float(fc["dist"]), np.array([0.0, 0.0, 1.0])))
            else:
                cands.append((0, C[i][2] - r, np.array([0.0, 0.0, 1.0])))
        for j, b in enumerate(boxes):
            sd, nrm = sd_box(C[i], np.asarray(b[:3], float), np.asarray(b[3:6], float))
            cands.append((off + j, sd - r, nrm))
        for j, sp in enumerate(spheres_obs):
            d = C[i] - np.asarray(sp[:3], float)
            dn = np.linalg.norm(d)
            cands.append((off + len(boxes) + j, dn - sp[3] - r, d / dn if dn > 0 else np.array([0, 0, 1.0])))
        for oi, d, nrm in cands:
            is_floor = floor and oi == 0
            if link >= 0 and d < best["o"]:  # arm_mindis.py:78 drops the base record
                best.update(o=d, arg_s=i, arg_o=oi, nrm=nrm)
            if not (is_floor and link in model["soft_floor_exempt"]):
                soft_min = min(soft_min, d)
            if not (is_floor and link in model["hard_floor_exempt"]):
                hard_min = min(hard_min, d)
    # self collision: basic_robot.py:89-98
    self_min = np.inf
    by_link = {}
    for i, s in enumerate(model["spheres"]):
        by_link.setdefault(s["link"], []).append(i)
    for a, b in model["self_pairs"]:
        for i in by_link[a]:
            for j in by_link[b]:
                d = np.linalg.norm(C[i] - C[j]) - model["spheres"][i]["r"] - model["spheres"][j]["r"]
                self_min = min(self_min, d)
    unsafe = hard_min <= thr_hard
    safe = (self_min > model["self_threshold"]) and (soft_min > thr_soft) and (not unsafe)
    # gradient: n^T J_lin(link, centre)[:, body_joints]   (arm_mindis.py:56-69)
    do_dq = np.zeros(n)
    if best["arg_s"] >= 0:
        link = model["spheres"][best["arg_s"]]["link"]
        p = C[best["arg_s"]]
        # ancestors of link
        anc = []
        l = link
        while l >= 0:
            anc.append(l)
            l = model["links"][l]["parent"]
        for l in anc:
            qi = model["links"][l]["qidx"]
            if qi >= 0:
                T = frames[l]
                z = T[:3, 2]
                do_dq[qi] = best["nrm"] @ np.cross(z, p - T[:3, 3])
    return dict(safe=bool(safe), unsafe=bool(unsafe), o=best["o"], arg_sphere=best["arg_s"], arg_obs=best["arg_o"],
                do_dq=do_dq, self_min=self_min, soft_min=soft_min, hard_min=hard_min)
