import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cbf-inc-rrt_b200"))
from cbfrrt import ops, workloads as W
dev = torch.device("cuda", 0)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
w = W.config(0, scale=0.25); n = w["n"]; rid = ops.ROBOT_ID[w["robot"]]
b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
net = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev)
common = (t(w["goal"]), b8, s4, 1, ops.no_grid(dev), net, t(w["K"]), t(w["u_lo"]), t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
q = t(w["q"])
for _ in range(6): ops.safety_filter(q, *common)
torch.cuda.synchronize()
print(q.shape)
