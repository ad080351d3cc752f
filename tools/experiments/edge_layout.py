"""cfg2: CTA-owns-whole-edges kernel (cbfrrt_edge_validity) vs one-thread-per-point over the concatenation (cbfrrt_edge_validity_var)"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cbf-inc-rrt_b200"))
from cbfrrt import ops, workloads as W
dev = torch.device("cuda", 0)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
w = W.config(2, scale=0.25); rid = ops.ROBOT_ID[w["robot"]]
b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
gr = ops.build_grid(b8, s4)
qf, qt = t(w["q"]), t(w["q_to"]); E = qf.shape[0]
K = torch.full((E,), 32, dtype=torch.int32)
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5
v1 = ops.edge_validity(qf, qt, 32, b8, s4, 1, gr, 0.02, 0.0, rid)
v2 = ops.edge_validity_var(qf, qt, K, b8, s4, 1, gr, 0.02, 0.0, rid)
print("equal:", torch.equal(v1[0], v2[0]), torch.equal(v1[1], v2[1]))
print("edge_validity     ms:", timeit(lambda: ops.edge_validity(qf, qt, 32, b8, s4, 1, gr, 0.02, 0.0, rid)))
print("edge_validity_var ms:", timeit(lambda: ops.edge_validity_var(qf, qt, K, b8, s4, 1, gr, 0.02, 0.0, rid)))
