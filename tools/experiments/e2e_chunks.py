"""e2e pipeline shape experiment: pinned q -> H2D -> fused kernel -> D2H (safe, u) with C chunks over S streams"""
import sys, os, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cbf-inc-rrt_b200"))
from cbfrrt import ops, workloads as W
dev = torch.device("cuda", 0)
w = W.config(1); n = w["n"]; rid = ops.ROBOT_ID[w["robot"]]
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
net = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev)
common = (t(w["goal"]), b8, s4, 1, ops.no_grid(dev), net, t(w["K"]), t(w["u_lo"]), t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
B = w["q"].shape[0]
qh = torch.from_numpy(w["q"]).pin_memory()
safe_h = torch.empty(B, dtype=torch.uint8).pin_memory(); u_h = torch.empty(B, n).pin_memory()
qd = torch.empty(B, n, device=dev)
def run(C, S):
    streams = [torch.cuda.Stream() for _ in range(S)]
    chunk = (B + C - 1) // C
    def once():
        for i in range(C):
            a, b = i * chunk, min(B, (i + 1) * chunk)
            with torch.cuda.stream(streams[i % S]):
                qd[a:b].copy_(qh[a:b], non_blocking=True)
                safe, u = ops.valid_control(qd[a:b], *common)
                safe_h[a:b].copy_(safe, non_blocking=True); u_h[a:b].copy_(u, non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(3): once()
    t0 = time.perf_counter()
    for _ in range(10): once()
    return (time.perf_counter() - t0) / 10
for C, S in ((1, 1), (2, 2), (3, 3), (4, 2), (4, 4), (6, 3), (8, 4), (12, 4), (16, 4)):
    dt = run(C, S)
    print(f"chunks {C:2d} streams {S}: {dt*1e3:.3f} ms  {B/dt/1e9:.2f}e9 states/s")
out = {"safe": safe_h.numpy(), "u": u_h.numpy()}
def host():
    ops.safety_filter_host(w["robot"], qh.numpy(), w["boxes"], w["spheres"], True, w["weights"], 48, 2, w["goal"], w["K"], w["u_lo"], w["u_hi"], w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], out, device=0)
for _ in range(3): host()
t0 = time.perf_counter()
for _ in range(10): host()
dt = (time.perf_counter() - t0) / 10
print(f"cbfrrt_safety_filter_host: {dt*1e3:.3f} ms  {B/dt/1e9:.2f}e9 states/s")
