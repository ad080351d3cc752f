#!/usr/bin/env python3
"""How far is the sphere model (spec/robot_models.json) from what the reference checks?

The reference asks Bullet about the CONVEX HULLS of the URDF collision meshes (neural_cbf/systems/arm_dynamics.py:169-187,
environment/basic_robot.py:89-98, environment/arm_env.py:82-84; pybullet loads mesh collision shapes as convex hulls).
This tool rebuilds those hulls from the meshes under /root/reference (dev container only), evaluates the exact
hull-vs-box / hull-vs-plane / hull-vs-hull distances (GJK, tools/geometry/gjk_hull.c, float64) on the reference's own
problem distribution (generate_problem_env.py:62-78: 8 boxes, pads / sticks / boxes, centres U([-.5,-.5,0],[.5,.5,1])),
and compares with the sphere model on the same states:

  * per link: over-approximation  hull distance - sphere-model distance  (positive = the spheres are fatter);
  * safe_mask / unsafe_mask confusion matrices of the two models (thresholds 0.02 / 0, the reference's exemptions);
  * self-collision pairs: hull-hull distance against the sphere model's, including the pairs the spec dropped.

    python tools/geometry/fidelity.py [--states 100000] [--scenes 25] [--spec spec/robot_models.json] [--out spec/geometry_fidelity]
Bullet's own margins (0.001 m default for URDF meshes) and its contact-breaking threshold are NOT modelled: both models
are compared at exactly the reference's numeric thresholds.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
from scipy.spatial import ConvexHull

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_robot_model as G  # noqa: E402


def gjk_lib():
    so = os.path.join(HERE, "_build", "libgjk.so")
    src = os.path.join(HERE, "gjk_hull.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm", "-pthread"])
    return C.CDLL(so)


def link_hulls(name):
    """pybullet link index (-1 .. n_links-1) -> hull vertices in the link frame (scaled, collision origin applied)"""
    if name == "panda":
        urdf, scale, root = f"{G.REF}/utils/robot/franka_panda/panda.urdf", 1.0, f"{G.REF}/utils/robot/franka_panda/"
    else:
        urdf, scale, root = f"{G.REF}/utils/robot/magician/urdf/demo.urdf", 2.0, f"{G.REF}/utils/robot/magician/"
    base, chain, coll = G.parse_urdf(urdf, scale)
    hulls, meshes = {}, {}
    for lidx in sorted(coll):
        mesh, cxyz, crpy = coll[lidx]
        rel = mesh.replace("package://magician/", "").replace("package://", "")
        if rel.endswith(".dae"):
            rel = rel[:-4] + ".STL"
        path = root + rel
        V, F = G.load_obj(path) if path.endswith(".obj") else G.load_stl(path)
        V = (V @ G.rpy_to_R(crpy).T + np.array(cxyz)) * scale
        hv = ConvexHull(V)
        hulls[lidx] = np.ascontiguousarray(V[hv.vertices])
        meshes[lidx] = (V, F, hv)
    return hulls, meshes


def fk_frames(spec, q):
    """float64 FK from the raw URDF numbers -> frames [B, n_links + 1, 12] (index 0 = base link -1): R (9) | p (3)"""
    links = spec["links"]
    B = q.shape[0]
    T = {-1: np.tile(np.eye(4), (B, 1, 1))}
    for l in links:
        M = np.eye(4)
        M[:3, :3] = G.rpy_to_R(l["rpy64"])
        M[:3, 3] = l["t64"]
        Tl = np.tile(M, (B, 1, 1))
        if l["qidx"] >= 0:
            a = q[:, l["qidx"]]
            Rz = np.tile(np.eye(4), (B, 1, 1))
            Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
            Tl = Tl @ Rz
        T[l["index"]] = T[l["parent"]] @ Tl
    out = np.zeros((B, len(links) + 1, 12))
    for i in range(-1, len(links)):
        out[:, i + 1, :9] = T[i][:, :3, :3].reshape(B, 9)
        out[:, i + 1, 9:] = T[i][:, :3, 3]
    return out


def reference_scene(rng, n_obs=8):
    """generate_problem_env.py:62-78"""
    sizes = np.zeros((n_obs, 3))
    for ii in range(n_obs):
        idx = rng.integers(0, 3)
        size = rng.uniform(0.05, 0.25, 3)
        if idx == 0:
            size[ii % 3] = 0.05
        elif idx == 1:
            size[ii % 3] = 0.05
            size[(ii + 1) % 3] = 0.05
        sizes[ii] = size
    while True:
        pos = rng.uniform((-0.5, -0.5, 0), (0.5, 0.5, 1), (n_obs, 3))
        far = [not (pos[i][2] - sizes[i][2] < 0.3 and (abs(abs(pos[i][0]) - sizes[i][0]) < 0.1 or abs(abs(pos[i][1]) - sizes[i][1]) < 0.1))
               for i in range(n_obs)]
        if all(far):
            return np.concatenate([pos, sizes], 1)


def sdf_box(c, boxes):
    """signed distance of points c [..., 3] to boxes [O, 6] -> [..., O]"""
    d = np.abs(c[..., None, :] - boxes[:, :3]) - boxes[:, 3:]
    return np.linalg.norm(np.maximum(d, 0), axis=-1) + np.minimum(d.max(-1), 0)


def analyse(name, spec, n_states, n_scenes, seed=0, threads=8):
    lib = gjk_lib()
    hulls, _ = link_hulls(name)
    n_links = spec["n_links"]
    nl1 = n_links + 1
    verts = np.concatenate([hulls.get(i, np.zeros((0, 3))) for i in range(-1, n_links)])
    voff = np.zeros(nl1 + 1, np.int32)
    voff[1:] = np.cumsum([len(hulls.get(i, ())) for i in range(-1, n_links)])
    rng = np.random.default_rng(seed)
    lo, hi = np.array(spec["q_lower"]), np.array(spec["q_upper"])
    sph_link = np.array([s["link"] for s in spec["spheres"]])
    sph_c = np.array([s["c"] for s in spec["spheres"]], np.float64)
    sph_r = np.array([s["r"] for s in spec["spheres"]], np.float64)
    pairs_ref = [tuple(p) for p in spec["self_pairs"]] + [tuple(p) for p in spec.get("self_pairs_dropped", [])]
    pairs_arr = np.array([(a + 1, b + 1) for a, b in pairs_ref], np.int32).reshape(-1, 2)
    per = n_states // n_scenes
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    acc = {"over": {l: [] for l in range(-1, n_links)}, "floor_over": {l: [] for l in range(-1, n_links)}}
    conf_safe, conf_unsafe = np.zeros((2, 2), np.int64), np.zeros((2, 2), np.int64)
    self_h, self_s = [], []
    for sc in range(n_scenes):
        boxes = reference_scene(rng)
        q = rng.uniform(lo, hi, (per, len(lo)))
        fr = fk_frames(spec, q)
        d_box = np.zeros((per, nl1, len(boxes)))
        d_self = np.zeros((per, len(pairs_ref)))
        lib.hull_distances(P(verts), P(voff), nl1, P(np.ascontiguousarray(fr)), C.c_long(per), P(np.ascontiguousarray(boxes)),
                           len(boxes), P(pairs_arr), len(pairs_ref), P(d_box), P(d_self), threads)
        # hull floor distance per link: lowest vertex
        d_floor_h = np.full((per, nl1), np.inf)
        for l in range(-1, n_links):
            V = hulls.get(l)
            if V is None:
                continue
            R3 = fr[:, l + 1, 6:9]                                   # third row of R
            d_floor_h[:, l + 1] = (V @ R3.T).min(0) + fr[:, l + 1, 11]
        # sphere model
        Rm = fr[:, sph_link + 1, :9].reshape(per, -1, 3, 3)
        cw = np.einsum("bsij,sj->bsi", Rm, sph_c) + fr[:, sph_link + 1, 9:]
        ds = sdf_box(cw, boxes) - sph_r[None, :, None]               # [B, S, O]
        d_box_s = np.full((per, nl1, len(boxes)), np.inf)
        d_floor_s = np.full((per, nl1), np.inf)
        for l in range(-1, n_links):
            m = sph_link == l
            if m.any():
                d_box_s[:, l + 1] = ds[:, m].min(1)
                d_floor_s[:, l + 1] = (cw[:, m, 2] - sph_r[m]).min(1)
        d_floor_s[:, 1] = spec["floor_const"]["dist"]                # link 0: kinematic constant
        for l in range(-1, n_links):
            if hulls.get(l) is None:
                continue
            h, s = d_box[:, l + 1], d_box_s[:, l + 1]
            m = (h > 0) & (h < 0.3)
            acc["over"][l].append((h - s)[m])
            m2 = (d_floor_h[:, l + 1] > 0) & (d_floor_h[:, l + 1] < 0.3)
            acc["floor_over"][l].append((d_floor_h[:, l + 1] - d_floor_s[:, l + 1])[m2])
        # self pairs (sphere model)
        ss = np.full((per, len(pairs_ref)), np.inf)
        for k, (a, b) in enumerate(pairs_ref):
            ia, ib = np.nonzero(sph_link == a)[0], np.nonzero(sph_link == b)[0]
            dd = np.linalg.norm(cw[:, ia][:, :, None] - cw[:, ib][:, None], axis=-1) - sph_r[ia][None, :, None] - sph_r[ib][None, None, :]
            ss[:, k] = dd.reshape(per, -1).min(1)
        self_h.append(d_self)
        self_s.append(ss)
        # masks of the two models (arm_dynamics.py:169-232, arm_env.py:82-84); kept pairs only for the sphere model
        n_keep = len(spec["self_pairs"])

        def masks(dbox, dfloor, dself, self_cols):
            soft_floor = dfloor.copy(); soft_floor[:, [0, 1]] = np.inf           # links -1, 0 exempt (arm_dynamics.py:180)
            hard_floor = dfloor.copy(); hard_floor[:, [0, 2]] = np.inf           # links -1, 1 filtered (arm_env.py:83-84)
            soft = np.minimum(dbox.min((1, 2)), soft_floor.min(1))
            hard = np.minimum(dbox.min((1, 2)), hard_floor.min(1))
            self_ok = (dself[:, self_cols] > 0.01).all(1) if len(self_cols) else np.ones(per, bool)
            unsafe = hard <= 0.0
            return self_ok & (soft > 0.02) & ~unsafe, unsafe
        safe_h, unsafe_h = masks(d_box, d_floor_h, d_self, list(range(len(pairs_ref))))   # reference: ALL its pairs
        safe_s, unsafe_s = masks(d_box_s, d_floor_s, ss, list(range(n_keep)))
        for a in (0, 1):
            for b in (0, 1):
                conf_safe[a, b] += int(((safe_h == a) & (safe_s == b)).sum())
                conf_unsafe[a, b] += int(((unsafe_h == a) & (unsafe_s == b)).sum())
    rep = {"robot": name, "states": per * n_scenes, "scenes": n_scenes, "spheres": int(len(sph_r)),
           "sphere_radius_range": [float(sph_r.min()), float(sph_r.max())], "links": {}}
    for l in range(-1, n_links):
        if hulls.get(l) is None:
            continue
        o = np.concatenate(acc["over"][l]) if acc["over"][l] else np.zeros(0)
        f = np.concatenate(acc["floor_over"][l]) if acc["floor_over"][l] else np.zeros(0)
        rep["links"][str(l)] = {"hull_vertices": int(len(hulls[l])), "spheres": int((sph_link == l).sum()),
                                "box_pairs": int(o.size), "over_mean": float(o.mean()) if o.size else None,
                                "over_p95": float(np.quantile(o, 0.95)) if o.size else None, "over_max": float(o.max()) if o.size else None,
                                "under_max": float(-o.min()) if o.size else None,
                                "floor_over_mean": float(f.mean()) if f.size else None, "floor_over_max": float(f.max()) if f.size else None,
                                "floor_under_max": float(-f.min()) if f.size else None}
    sh, ss_ = np.concatenate(self_h), np.concatenate(self_s)
    rep["self_pairs"] = {}
    for k, (a, b) in enumerate(pairs_ref):
        rep["self_pairs"][f"{a}-{b}"] = {"kept": k < len(spec["self_pairs"]), "hull_min": float(sh[:, k].min()),
                                         "hull_within_0.01": float((sh[:, k] <= 0.01).mean()),
                                         "spheres_within_0.01": float((ss_[:, k] <= 0.01).mean()),
                                         "over_mean": float((sh[:, k] - ss_[:, k])[sh[:, k] > 0].mean())}
    tot = conf_safe.sum()
    rep["safe_mask"] = {"hull_safe_rate": float(conf_safe[1].sum() / tot), "model_safe_rate": float(conf_safe[:, 1].sum() / tot),
                        "false_unsafe_of_hull_safe": float(conf_safe[1, 0] / max(conf_safe[1].sum(), 1)),
                        "false_safe_of_hull_not_safe": float(conf_safe[0, 1] / max(conf_safe[0].sum(), 1)),
                        "confusion[hull][model]": conf_safe.tolist()}
    rep["unsafe_mask"] = {"hull_unsafe_rate": float(conf_unsafe[1].sum() / tot), "model_unsafe_rate": float(conf_unsafe[:, 1].sum() / tot),
                          "missed_collisions_of_hull_unsafe": float(conf_unsafe[1, 0] / max(conf_unsafe[1].sum(), 1)),
                          "false_collisions_of_hull_free": float(conf_unsafe[0, 1] / max(conf_unsafe[0].sum(), 1)),
                          "confusion[hull][model]": conf_unsafe.tolist()}
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=100000)
    ap.add_argument("--scenes", type=int, default=25)
    ap.add_argument("--spec", default=os.path.join(ROOT, "spec", "robot_models.json"))
    ap.add_argument("--out", default=os.path.join(ROOT, "spec", "geometry_fidelity"))
    ap.add_argument("--robots", default="panda,magician")
    a = ap.parse_args()
    models = json.load(open(a.spec))
    out = {}
    for name in a.robots.split(","):
        out[name] = analyse(name, models[name], a.states, a.scenes)
        r = out[name]
        print(name, "safe_mask", {k: v for k, v in r["safe_mask"].items() if not k.startswith("conf")})
        print(name, "unsafe_mask", {k: v for k, v in r["unsafe_mask"].items() if not k.startswith("conf")})
        for l, v in r["links"].items():
            print(f"  link {l:>2}: spheres {v['spheres']}, over mean {v['over_mean']:.4f} p95 {v['over_p95']:.4f} max {v['over_max']:.4f} under max {v['under_max']:.4f}")
        for p, v in r["self_pairs"].items():
            print(f"  self {p}: kept {v['kept']} hull<=0.01 {v['hull_within_0.01']:.4f} spheres<=0.01 {v['spheres_within_0.01']:.4f} hull_min {v['hull_min']:.4f}")
    json.dump(out, open(a.out + ".json", "w"), indent=1)


if __name__ == "__main__":
    main()
