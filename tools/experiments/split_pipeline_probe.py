#!/usr/bin/env python3
"""fused safety kernel vs a two-pass pipeline (geometry pass -> datax in L2-sized chunks -> barrier-network pass) on the
bench scene (Panda, 256 boxes): is the single fused launch still the fastest form?  One JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from cbfrrt import ops, workloads as W  # noqa: E402

dev = "cuda"
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
out = {}
for cfg, robot in ((4, "panda"), (1, "magician")):
    w = W.config(cfg, scale=1e-9)
    n, rid = w["n"], ops.ROBOT_ID[robot]
    B = 1 << 23 if cfg == 4 else 1 << 20
    b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
    gr = ops.build_grid(b8, s4)
    net = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev)
    q = ops.sample_states(B, 0, 0, rid, device=dev)
    K, lo, hi, goal = t(w["K"]), t(w["u_lo"]), t(w["u_hi"]), t(w["goal"][:1])
    empty = torch.empty(0, device=dev)

    def fused():
        return ops.valid_control(q, goal, b8, s4, 1, gr, net, K, lo, hi, w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)

    def split(chunk):
        us, vs = [], []
        for o in range(0, B, chunk):
            s_, u_, d_ = ops.observe(q[o:o + chunk], b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid)
            h, J, u, r = ops.cbf_filter(d_, goal, net, K, lo, hi, w["lam"], w["penalty"], 48, 2, empty)
            us.append(u); vs.append(s_)
        return vs, us

    def timed(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(iters):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    v0, u0 = fused()
    vs, us = split(1 << 20)
    same = bool(torch.equal(torch.cat(vs), v0) and torch.equal(torch.cat(us), u0))
    res = {"states": B, "fused_ms": timed(fused), "bitwise_equal": same}
    for ch in (1 << 18, 1 << 19, 1 << 20, 1 << 21, B):
        if ch <= B:
            res[f"split_chunk_{ch}_ms"] = timed(lambda: split(ch))
    out[f"cfg{cfg}_{robot}"] = res
print(json.dumps(out))
