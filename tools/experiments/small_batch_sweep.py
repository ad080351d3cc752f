"""FP32-FMA vs tcgen05 kernels as a function of the batch size (run twice: CBFRRT_MLP=fma and CBFRRT_MLP=tc)"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cbf-inc-rrt_b200"))
from cbfrrt import ops, workloads as W
dev = torch.device("cuda", 0)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
for cfg in (0, 1):
    w = W.config(cfg, scale=0.001); n = w["n"]; rid = ops.ROBOT_ID[w["robot"]]
    b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
    net = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev)
    common = (t(w["goal"]), b8, s4, 1, ops.no_grid(dev), net, t(w["K"]), t(w["u_lo"]), t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
    for B in (1024, 4096, 8192, 16384, 32768, 65536, 131072):
        q = t(W.sample_states(w["robot"], B, 0))
        for _ in range(5): ops.safety_filter(q, *common)
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.safety_filter(q, *common); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(os.environ.get("CBFRRT_MLP", "auto"), w["robot"], B, f"{np.median(ts)*1e3:.1f} us")
