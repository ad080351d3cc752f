#!/usr/bin/env python3
"""one warm-up + a few planner-sized rollout calls (T trajectories x 200 steps, Panda, 8 boxes) -- the command ncu wraps"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from cbfrrt import ops, workloads as W  # noqa: E402

T = int(os.environ.get("T", "64"))
dev = "cuda"
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
w = W.config(3, scale=1e-9)
boxes = W.scene(3)[0]
b8, s4 = ops.pack_scene(boxes[:, :3], boxes[:, 3:], None, device=dev)
gr = ops.build_grid(b8, s4)
net = ops.pack_weights(w["state_dict"], 7, 48, 2, device=dev)
cand = W.sample_states("panda", 40000, 7)
pool = cand[ops.masks(t(cand), b8, s4, 1, gr, 0.02, 0.0, 0)[0].bool().cpu().numpy()]
q0, goal = t(np.resize(pool, (T, 7))), t(W.sample_states("panda", T, 8))
for _ in range(4):
    out = ops.rollout(q0, goal, 200, 4, 1 / 120, 9, 10, 3e-2, True, False, b8, s4, 1, gr, net, t(w["K"]), t(w["u_lo"]), t(w["u_hi"]),
                      w["lam"], w["penalty"], 0.02, 0.0, 0, 48, 2)
    q64 = t(W.sample_states("panda", 64, 3))
    ops.safety_filter(q64, goal[:1], b8, s4, 1, gr, net, t(w["K"]), t(w["u_lo"]), t(w["u_hi"]), w["lam"], w["penalty"], 0.02, 0.0, 0, 48, 2)
torch.cuda.synchronize()
print("steps done (mean):", float(out[2].float().mean()))
