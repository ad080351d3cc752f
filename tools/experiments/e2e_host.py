"""time cbfrrt_safety_filter_host on cfg1 (pinned host buffers); chunking via CBFRRT_HOST_CHUNKS"""
import sys, os, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cbf-inc-rrt_b200"))
from cbfrrt import ops, workloads as W
w = W.config(1); n = w["n"]; B = w["q"].shape[0]
qh = torch.from_numpy(w["q"]).pin_memory()
safe_h = torch.empty(B, dtype=torch.uint8).pin_memory(); u_h = torch.empty(B, n).pin_memory()
out = {"safe": safe_h.numpy(), "u": u_h.numpy()}
def host():
    ops.safety_filter_host(w["robot"], qh.numpy(), w["boxes"], w["spheres"], True, w["weights"], 48, 2, w["goal"], w["K"], w["u_lo"], w["u_hi"], w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], out, device=0)
for _ in range(5): host()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); host(); ts.append(time.perf_counter() - t0)
dt = float(np.median(ts))
print(f"chunks={os.environ.get('CBFRRT_HOST_CHUNKS','default')}: {dt*1e3:.3f} ms  {B/dt/1e9:.2f}e9 states/s (min {min(ts)*1e3:.3f} ms)")
