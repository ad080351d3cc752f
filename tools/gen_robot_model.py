#!/usr/bin/env python3
"""Spec freeze: derive the robot model tables (FK constants + link sphere geometry).

Runs ONLY in the dev container (reads the reference's URDF + collision meshes as a
DATA source; /root/reference does not exist on the GPU box).  The output is committed:

  spec/robot_models.json                     single source of truth (float32-exact decimals)
  cbf-inc-rrt_b200/csrc/robot_tables.cuh     constexpr tables for the CUDA kernels
  oracle/robot_tables.h                      the same numbers for the C oracle
  cbf-inc-rrt_b200/cbfrrt/_robot_tables.py   the same numbers for the Python host layer

What is derived (reference file:line it follows)
  * chain of pybullet "links" (= URDF joints, in URDF order; pybullet link i is the child of
    joint i): parent, joint origin xyz/rpy, joint type      utils/robot/franka_panda/panda.urdf:47-336
                                                            utils/robot/magician/urdf/demo.urdf:38-187
  * globalScaling=2 for the Magician                        environment/magician.py:20
  * joint limits / velocity limits / q0                     panda.urdf <limit>, environment/franka_panda.py:23-24,
                                                            environment/magician.py:22
  * finger prismatic joints frozen at 0 (one POSITION_CONTROL step only, never reset afterwards)
                                                            environment/basic_robot.py:106-110
  * link geometry: Bullet uses convex hulls of the collision meshes; that is not reproducible
    without pybullet and is a poor GPU shape.  We DEFINE the geometry as a per-link set of spheres
    fitted deterministically to the collision-mesh surface (every surface sample, <= 6 mm spacing,
    lies inside at least one sphere of its link).  SURVEY.md section 7 hard-part 1.

Deterministic fixed-rotation contract: the joint-origin rotations are computed in float64 from the
URDF rpy, entries with |v| < 1e-7 are snapped to 0 and |v|-1 < 1e-7 to +-1 (rpy = 1.57079632679 is
not exactly pi/2), then rounded to float32.
"""
import json
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = os.environ.get("CBFRRT_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------- mesh IO
def load_obj(path):
    V, F = [], []
    for line in open(path):
        if line.startswith("v "):
            V.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("f "):
            idx = [int(t.split("/")[0]) - 1 for t in line.split()[1:]]
            for k in range(1, len(idx) - 1):
                F.append([idx[0], idx[k], idx[k + 1]])
    return np.array(V, dtype=np.float64), np.array(F, dtype=np.int64)


def load_stl(path):
    b = open(path, "rb").read()
    n = struct.unpack("<I", b[80:84])[0]
    assert len(b) == 84 + 50 * n, "binary STL expected"
    rec = np.frombuffer(b[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    V = rec["v"].reshape(-1, 3).astype(np.float64)
    F = np.arange(V.shape[0]).reshape(-1, 3)
    return V, F


def surface_samples(V, F, max_edge):
    """vertices + barycentric grid samples of every triangle so that spacing <= max_edge."""
    pts = [V]
    tri = V[F]  # (T,3,3)
    e = np.stack([np.linalg.norm(tri[:, 0] - tri[:, 1], axis=1),
                  np.linalg.norm(tri[:, 1] - tri[:, 2], axis=1),
                  np.linalg.norm(tri[:, 2] - tri[:, 0], axis=1)], axis=1).max(axis=1)
    nsub = np.ceil(e / max_edge).astype(int)
    for n in np.unique(nsub):
        if n <= 1:
            continue
        sel = tri[nsub == n]
        ij = [(i, j) for i in range(n + 1) for j in range(n + 1 - i)]
        w = np.array([[i, j, n - i - j] for i, j in ij], dtype=np.float64) / n  # (S,3)
        pts.append(np.einsum("sk,tkd->tsd", w, sel).reshape(-1, 3))
    return np.concatenate(pts, axis=0)


# ----------------------------------------------------------------------------- sphere fit
def bounding_sphere(P, iters=400):
    """Badoiu-Clarkson core-set iteration (deterministic), then exact covering radius."""
    c = P.mean(axis=0)
    for i in range(1, iters + 1):
        d = np.linalg.norm(P - c, axis=1)
        far = P[int(np.argmax(d))]
        c = c + (far - c) / (i + 1)
    r = float(np.linalg.norm(P - c, axis=1).max())
    return c, r


def fit_spheres(P, k, iters=40):
    """k spheres covering P: deterministic k-centre clustering.

    farthest-point initialisation (first seed = sample farthest from the centroid), then Lloyd-style
    rounds where every cluster centre moves to the centre of its cluster's bounding sphere; the final
    assignment is made on ALL samples and each cluster gets its exact covering radius.
    """
    Q = P[::max(1, len(P) // 6000)]
    seeds = [Q[int(np.argmax(np.linalg.norm(Q - Q.mean(axis=0), axis=1)))]]
    d = np.linalg.norm(Q - seeds[0], axis=1)
    for _ in range(1, k):
        seeds.append(Q[int(np.argmax(d))])
        d = np.minimum(d, np.linalg.norm(Q - seeds[-1], axis=1))
    C = np.array(seeds)
    for _ in range(iters):
        a = np.argmin(np.linalg.norm(Q[:, None] - C[None], axis=2), axis=1)
        newC = C.copy()
        for j in range(k):
            if (a == j).any():
                newC[j], _ = bounding_sphere(Q[a == j], iters=100)
        if np.allclose(newC, C, atol=1e-6):
            break
        C = newC
    a = np.argmin(np.linalg.norm(P[:, None] - C[None], axis=2), axis=1)
    out = [bounding_sphere(P[a == j]) for j in range(k) if (a == j).any()]
    # stable order: along the principal axis of the link
    mu = P.mean(axis=0)
    axis = np.linalg.svd(P - mu, full_matrices=False)[2][0]
    j = int(np.argmax(np.abs(axis)))
    if axis[j] < 0:
        axis = -axis
    out.sort(key=lambda cr: float((cr[0] - mu) @ axis))
    return out


# ----------------------------------------------------------------------------- URDF
def rpy_to_R(rpy):
    r, p, y = rpy
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def snap(R):
    R = R.copy()
    R[np.abs(R) < 1e-7] = 0.0
    R[np.abs(R - 1) < 1e-7] = 1.0
    R[np.abs(R + 1) < 1e-7] = -1.0
    return R


def parse_urdf(path, scale):
    root = ET.parse(path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    child_names = {j.find("child").get("link") for j in joints}
    base = [n for n in links if n not in child_names]
    assert len(base) == 1
    base = base[0]
    name_to_idx = {base: -1}
    chain = []
    for i, j in enumerate(joints):
        o = j.find("origin")
        xyz = [float(x) for x in o.get("xyz").split()]
        rpy = [float(x) for x in o.get("rpy").split()]
        axis = j.find("axis")
        ax = [float(x) for x in axis.get("xyz").split()] if axis is not None else [0, 0, 0]
        lim = j.find("limit")
        child = j.find("child").get("link")
        name_to_idx[child] = i
        chain.append(dict(
            joint=j.get("name"), type=j.get("type"), link=child,
            parent=name_to_idx[j.find("parent").get("link")],
            xyz=(np.array(xyz) * scale).tolist(), rpy=rpy, axis=ax,
            lower=float(lim.get("lower")) if lim is not None else 0.0,
            upper=float(lim.get("upper")) if lim is not None else 0.0,
            velocity=float(lim.get("velocity")) if lim is not None else 0.0,
        ))
    # collision meshes (URDF collision origin; only the right finger has one: yaw pi)
    coll = {}
    for name, l in links.items():
        c = l.find("collision")
        if c is None:
            continue
        mesh = c.find("geometry").find("mesh").get("filename")
        o = c.find("origin")
        cxyz = [float(x) for x in o.get("xyz").split()] if o is not None else [0, 0, 0]
        crpy = [float(x) for x in o.get("rpy").split()] if o is not None else [0, 0, 0]
        coll[name_to_idx[name]] = (mesh, cxyz, crpy)
    return base, chain, coll


def f32(x):
    return float(np.float32(x))


def build_robot(name):
    if name == "panda":
        urdf = f"{REF}/utils/robot/franka_panda/panda.urdf"
        scale = 1.0
        mesh_root = f"{REF}/utils/robot/franka_panda/"
        q0 = [1.00887519, 0.50546576, -1.69052917, -2.2909179, 2.95208592, 3.59793418, 2.93001438]
        # spheres per link (pybullet link index): chosen once, frozen here
        n_spheres = {-1: 4, 0: 3, 1: 3, 2: 3, 3: 3, 4: 5, 5: 3, 6: 2, 8: 4, 9: 1, 10: 1}
    elif name == "magician":
        urdf = f"{REF}/utils/robot/magician/urdf/demo.urdf"
        scale = 2.0  # environment/magician.py:20 globalScaling=2
        mesh_root = f"{REF}/utils/robot/magician/"
        q0 = [0, 0.707, 0.707, 0]
        n_spheres = {-1: 4, 0: 7, 1: 5, 2: 5, 3: 3}
    else:
        raise ValueError(name)
    base, chain, coll = parse_urdf(urdf, scale)

    # ---- chain tables
    links = []
    body_joints = []
    for i, c in enumerate(chain):
        R = snap(rpy_to_R(c["rpy"]))
        if c["type"] == "revolute":
            assert c["axis"] == [0, 0, 1], "all revolute axes are +z in both URDFs"
            body_joints.append(i)
            qidx = len(body_joints) - 1
        else:
            qidx = -1  # fixed, or prismatic frozen at displacement 0
        links.append(dict(
            index=i, name=c["link"], joint=c["joint"], type=c["type"], parent=c["parent"], qidx=qidx,
            t=[f32(v) for v in c["xyz"]], R=[[f32(v) for v in row] for row in R],
            t64=c["xyz"], rpy64=c["rpy"],
            lower=c["lower"], upper=c["upper"], velocity=c["velocity"],
        ))
    n = len(body_joints)

    # ---- spheres
    spheres = []
    for lidx in sorted(coll):
        mesh, cxyz, crpy = coll[lidx]
        rel = mesh.replace("package://magician/", "").replace("package://", "")
        if rel.endswith(".dae"):
            # demo.urdf:170 points link_5's collision at the .dae; its vertex AABB equals link_5.STL's
            # (checked when this script was written), and the STL needs no COLLADA parser.
            rel = rel[:-4] + ".STL"
        path = mesh_root + rel
        V, F = load_obj(path) if path.endswith(".obj") else load_stl(path)
        Rc = rpy_to_R(crpy)
        V = (V @ Rc.T + np.array(cxyz)) * scale
        P = surface_samples(V, F, max_edge=0.006 * scale)
        if lidx == 0:
            link0_min_z = float(V[:, 2].min())
        fits = fit_spheres(P, n_spheres[lidx])
        for c, r in fits:
            r = r + 1e-4  # pad so that fp32 rounding of c cannot uncover a sample
            spheres.append(dict(link=lidx, c=[f32(v) for v in c], r=f32(r)))
        # coverage check with the float32 numbers
        C = np.array([s["c"] for s in spheres if s["link"] == lidx])
        Rr = np.array([s["r"] for s in spheres if s["link"] == lidx])
        d = np.linalg.norm(P[:, None, :] - C[None], axis=2) - Rr[None]
        assert (d.min(axis=1) <= 0).all(), f"{name} link {lidx}: sample not covered"

    # ---- link 0 vs floor is a kinematic constant (link 0 only spins about the world z axis): the covering
    # spheres of a flat-bottomed housing would dip through the floor, so this one pair uses the exact
    # convex-hull distance  (joint z offset + lowest mesh vertex)  instead of the spheres.
    l0 = links[0]
    assert l0["parent"] == -1 and l0["qidx"] == 0 and l0["R"][2][2] == 1.0
    floor_const = dict(link=0, dist=f32(l0["t64"][2] + link0_min_z))

    # ---- exemptions / pair lists
    n_links = len(links)
    # arm_dynamics.py:180: links -1 and 0 are exempt from the floor in the soft check
    soft_floor_exempt = [-1, 0]
    # arm_env.py:83-84: collision filter robot link -1 and link 1 vs plane (hard/contact check)
    hard_floor_exempt = [-1, 1]
    # basic_robot.py:89-98
    self_pairs = []
    for a in range(n_links):
        for bi in range(a + 2, n):
            if n == 7 and a == 6 and bi == 8:
                continue
            self_pairs.append([a, body_joints[bi]])
    has_geom = sorted(set(s["link"] for s in spheres))
    self_pairs = [p for p in self_pairs if p[0] in has_geom and p[1] in has_geom]
    # Geometry-model artefact rule: the covering spheres are ~1-2 cm fatter than the meshes, so a pair of
    # links that is structurally a few mm apart for EVERY configuration (Panda link5/link7 around the
    # wrist) would report "self-collision" always and make safe_mask identically False.  A pair is
    # dropped when the sphere model has it within the 0.01 m threshold in more than half of 512 uniformly
    # sampled configurations (fixed seed, float64 FK).  The dropped pairs are recorded in the spec.
    dropped = []
    rng = np.random.default_rng(12345)
    lo = np.array([links[j]["lower"] for j in body_joints])
    hi = np.array([links[j]["upper"] for j in body_joints])
    hits = {tuple(p): 0 for p in self_pairs}
    n_samp = 512
    by_link = {}
    for i, s in enumerate(spheres):
        by_link.setdefault(s["link"], []).append(i)
    for _ in range(n_samp):
        q = rng.uniform(lo, hi)
        frames = {-1: np.eye(4)}
        for l in links:
            T = np.eye(4)
            T[:3, :3] = rpy_to_R(l["rpy64"])
            T[:3, 3] = l["t64"]
            if l["qidx"] >= 0:
                a = q[l["qidx"]]
                Rz = np.eye(4)
                Rz[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
                T = T @ Rz
            frames[l["index"]] = frames[l["parent"]] @ T
        C = np.array([frames[s["link"]][:3, :3] @ np.array(s["c"]) + frames[s["link"]][:3, 3] for s in spheres])
        R = np.array([s["r"] for s in spheres])
        for a, b in self_pairs:
            ia, ib = by_link[a], by_link[b]
            d = np.linalg.norm(C[ia][:, None] - C[ib][None], axis=2) - R[ia][:, None] - R[ib][None]
            if d.min() <= 0.01:
                hits[(a, b)] += 1
    for p in list(self_pairs):
        if hits[tuple(p)] > n_samp // 2:
            dropped.append(p)
            self_pairs.remove(p)
    print(name, "self-pair hit rates", {k: round(v / n_samp, 3) for k, v in hits.items()}, "dropped", dropped)

    # ---- derived kernel tables
    # spheres are sorted by link; sph_begin[l+1] .. sph_begin[l+2] is the range of link l (l = -1 .. n_links-1)
    sph_begin = [0]
    for l in range(-1, n_links):
        sph_begin.append(sph_begin[-1] + sum(1 for s in spheres if s["link"] == l))
    assert [s["link"] for s in spheres] == sorted(s["link"] for s in spheres)
    # bit j set <=> body joint j moves link l
    anc_mask = []
    for l in range(n_links):
        m, k = 0, l
        while k >= 0:
            if links[k]["qidx"] >= 0:
                m |= 1 << links[k]["qidx"]
            k = links[k]["parent"]
        anc_mask.append(m)
    # exact squared-distance thresholds for the self-collision test: the oracle's predicate
    #   fl(fl(sqrtf(dsq) - r_i) - r_j) > 0.01f        (monotone non-decreasing in dsq)
    # is equivalent to  dsq > X_ij  with  X_ij = the largest float32 for which the predicate is false.
    def pred_false(x, ri, rj):
        d = np.float32(np.float32(np.sqrt(np.float32(x))) - np.float32(ri)) - np.float32(rj)
        return not (np.float32(d) > np.float32(0.01))
    self_sphere_pairs = []
    for a, b in self_pairs:
        for i in by_link[a]:
            for j in by_link[b]:
                ri, rj = spheres[i]["r"], spheres[j]["r"]
                lo_b, hi_b = 0, int(np.float32(4.0).view(np.uint32))
                assert pred_false(np.uint32(lo_b).view(np.float32), ri, rj)
                assert not pred_false(np.uint32(hi_b).view(np.float32), ri, rj)
                while hi_b - lo_b > 1:
                    mid = (lo_b + hi_b) // 2
                    if pred_false(np.uint32(mid).view(np.float32), ri, rj):
                        lo_b = mid
                    else:
                        hi_b = mid
                self_sphere_pairs.append([i, j, float(np.uint32(lo_b).view(np.float32))])

    return dict(
        sph_begin=sph_begin, anc_mask=anc_mask, self_sphere_pairs=self_sphere_pairs,
        name=name, scale=scale, n=n, n_links=n_links, body_joints=body_joints,
        ee_joints=[i for i, l in enumerate(links) if l["type"] == "prismatic"],
        q0=q0,
        q_lower=[links[j]["lower"] for j in body_joints],
        q_upper=[links[j]["upper"] for j in body_joints],
        v_max=[links[j]["velocity"] for j in body_joints],
        base_link=base, links=links, spheres=spheres, floor_const=floor_const,
        soft_floor_exempt=soft_floor_exempt, hard_floor_exempt=hard_floor_exempt,
        self_pairs=self_pairs, self_pairs_dropped=dropped, self_threshold=0.01,
    )


def link_balls(spheres, n_links):
    """per link (-1 .. n_links-1) a ball [cx, cy, cz, R] in the link frame that contains all of the link's spheres
    (Badoiu-Clarkson iteration towards the farthest sphere; R padded).  Used by the kernels' exact culling test."""
    balls = []
    for l in range(-1, n_links):
        P = np.array([s["c"] for s in spheres if s["link"] == l], dtype=np.float64)
        r = np.array([s["r"] for s in spheres if s["link"] == l], dtype=np.float64)
        if len(P) == 0:
            balls.append([0.0, 0.0, 0.0, 0.0])
            continue
        c = P.mean(0)
        for it in range(1, 4001):
            i = int(np.argmax(np.linalg.norm(P - c, axis=1) + r))
            c = c + (P[i] - c) / (it + 1)
        c32 = np.float32(c).astype(np.float64)
        rad = float((np.linalg.norm(P - c32, axis=1) + r).max() * (1 + 1e-6) + 1e-6)
        balls.append([float(c32[0]), float(c32[1]), float(c32[2]), rad])
    return balls


# ----------------------------------------------------------------------------- emitters
def cfloat(x):
    s = repr(float(np.float32(x)))
    if "e" not in s and "." not in s and "inf" not in s:
        s += ".0"
    return s + "f"


def emit_c_tables(robots, path, cuda):
    q = "__device__ constexpr" if cuda else "static const"
    L = []
    L.append("// GENERATED by tools/gen_robot_model.py from spec/robot_models.json -- do not edit.")
    L.append("// Robot model tables: FK chain constants + link sphere geometry (see DESIGN.md, 'Geometry model').")
    L.append("#pragma once")
    for rb in robots:
        N = rb["name"].upper()
        nl, ns, n = rb["n_links"], len(rb["spheres"]), rb["n"]
        L.append(f"// ---- {rb['name']}: n={n} links={nl} spheres={ns}")
        L.append(f"#define {N}_N {n}")
        L.append(f"#define {N}_NLINKS {nl}")
        L.append(f"#define {N}_NSPHERES {ns}")
        L.append(f"#define {N}_NSELFPAIRS {len(rb['self_pairs'])}")
        L.append(f"#define {N}_LINK0_FLOOR_DIST {cfloat(rb['floor_const']['dist'])}")
        L.append(f"{q} int {N}_PARENT[{nl}] = {{" + ", ".join(str(l["parent"]) for l in rb["links"]) + "};")
        L.append(f"{q} int {N}_QIDX[{nl}] = {{" + ", ".join(str(l["qidx"]) for l in rb["links"]) + "};")
        L.append(f"{q} float {N}_T[{nl}][3] = {{")
        for l in rb["links"]:
            L.append("  {" + ", ".join(cfloat(v) for v in l["t"]) + "},")
        L.append("};")
        L.append(f"{q} float {N}_R[{nl}][9] = {{")
        for l in rb["links"]:
            L.append("  {" + ", ".join(cfloat(v) for row in l["R"] for v in row) + "},")
        L.append("};")
        L.append(f"{q} int {N}_SPH_LINK[{ns}] = {{" + ", ".join(str(s["link"]) for s in rb["spheres"]) + "};")
        L.append(f"{q} float {N}_SPH[{ns}][4] = {{")
        for s in rb["spheres"]:
            L.append("  {" + ", ".join(cfloat(v) for v in s["c"] + [s["r"]]) + "},")
        L.append("};")
        L.append(f"{q} int {N}_SELF_PAIRS[{max(1, len(rb['self_pairs']))}][2] = {{" +
                 ", ".join("{%d, %d}" % tuple(p) for p in rb["self_pairs"]) + "};")
        if cuda:
            L.append(f"{q} int {N}_SPH_BEGIN[{nl + 2}] = {{" + ", ".join(str(v) for v in rb["sph_begin"]) + "};")
            L.append(f"{q} int {N}_ANC_MASK[{nl}] = {{" + ", ".join(str(v) for v in rb["anc_mask"]) + "};")
            L.append(f"{q} float {N}_LINK_BALL[{nl + 1}][4] = {{")
            for bl in rb["link_balls"]:
                L.append("  {" + ", ".join(cfloat(v) for v in bl) + "},")
            L.append("};")
            sp = rb["self_sphere_pairs"]
            L.append(f"#define {N}_NSELFSPH {len(sp)}")
            L.append(f"{q} int {N}_SELFSPH_I[{len(sp)}] = {{" + ", ".join(str(v[0]) for v in sp) + "};")
            L.append(f"{q} int {N}_SELFSPH_J[{len(sp)}] = {{" + ", ".join(str(v[1]) for v in sp) + "};")
            L.append(f"{q} float {N}_SELFSPH_X[{len(sp)}] = {{" + ", ".join(cfloat(v[2]) for v in sp) + "};")
        L.append(f"{q} float {N}_QLO[{n}] = {{" + ", ".join(cfloat(v) for v in rb["q_lower"]) + "};")
        L.append(f"{q} float {N}_QHI[{n}] = {{" + ", ".join(cfloat(v) for v in rb["q_upper"]) + "};")
        L.append(f"{q} float {N}_VMAX[{n}] = {{" + ", ".join(cfloat(v) for v in rb["v_max"]) + "};")
        L.append("")
    open(path, "w").write("\n".join(L) + "\n")


def emit_py_tables(robots, path):
    slim = []
    for rb in robots:
        r = {k: v for k, v in rb.items()}
        slim.append(r)
    with open(path, "w") as f:
        f.write('"""GENERATED by tools/gen_robot_model.py from spec/robot_models.json -- do not edit."""\n')
        f.write("ROBOT_MODELS = ")
        f.write(json.dumps({r["name"]: r for r in slim}, indent=1))
        f.write("\n")


def main():
    if "--from-spec" in sys.argv:     # re-derive the kernel tables from the committed spec (no meshes needed)
        with open(f"{ROOT}/spec/robot_models.json") as f:
            spec = json.load(f)
        robots = [spec["panda"], spec["magician"]]
    else:
        robots = [build_robot("panda"), build_robot("magician")]
    for rb in robots:
        rb["link_balls"] = link_balls(rb["spheres"], rb["n_links"])
    os.makedirs(f"{ROOT}/spec", exist_ok=True)
    with open(f"{ROOT}/spec/robot_models.json", "w") as f:
        json.dump({r["name"]: r for r in robots}, f, indent=1)
    emit_c_tables(robots, f"{ROOT}/cbf-inc-rrt_b200/csrc/robot_tables.cuh", cuda=True)
    emit_c_tables(robots, f"{ROOT}/oracle/robot_tables.h", cuda=False)
    emit_py_tables(robots, f"{ROOT}/cbf-inc-rrt_b200/cbfrrt/_robot_tables.py")
    for rb in robots:
        print(rb["name"], "n", rb["n"], "links", rb["n_links"], "spheres", len(rb["spheres"]),
              "self pairs", rb["self_pairs"])
        for s in rb["spheres"]:
            print("  link %2d c=(%.4f %.4f %.4f) r=%.4f" % (s["link"], *s["c"], s["r"]))


if __name__ == "__main__":
    sys.exit(main())
