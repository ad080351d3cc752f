#!/usr/bin/env python3
"""Device-resident timings of every BASELINE.json config (1 GPU), for BASELINE.md / DESIGN.md tables.

Not the driver's benchmark (that is bench.py, configs[1]); this prints one JSON line per config with CUDA-event
timings (>= 3 warm-ups, L2 flushed between iterations) and, for configs small enough, the CPU port beside it.
    python tools/measure_configs.py [--scale S] [--iters K] [--only cfg0,cfg2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
from cbfrrt import ops, workloads as W  # noqa: E402


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
        flush.zero_()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
        flush.zero_()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="cfg0,cfg1,cfg2,cfg3,cfg4")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU port on a bounded sample")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    threads = os.cpu_count() or 1
    if args.cpu:
        from oracle import c_oracle
        c_oracle.set_threads(threads)
    for name in args.only.split(","):
        idx = int(name[3:])
        scale = args.scale * (1 / 8 if idx == 4 else 1.0)   # cfg4: 2^23 states per launch, reported per state
        w = W.config(idx, scale=scale)
        rid, n = ops.ROBOT_ID[w["robot"]], w["n"]
        b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
        gr = ops.build_grid(b8, s4) if b8.shape[0] + s4.shape[0] > 0 else ops.no_grid(dev)
        net = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev)
        K, lo, hi = t(w["K"]), t(w["u_lo"]), t(w["u_hi"])
        goal = t(w["goal"])
        res = {"config": w["name"], "robot": w["robot"], "boxes": int(w["boxes"].shape[0]), "spheres": int(w["spheres"].shape[0])}
        cpu_fn = None
        if idx in (0, 1, 4):
            q = t(w["q"]) if w["q"] is not None else ops.sample_states(w["n_states"], 0, 0, rid, device=dev)
            B = q.shape[0]
            common = (goal, b8, s4, 1, gr, net, K, lo, hi, w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
            variants = {"safety_filter(all outputs)": lambda: ops.safety_filter(q, *common),
                        "valid_control(valid,u)": lambda: ops.valid_control(q, *common),
                        "observe(masks+datax)": lambda: ops.observe(q, b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid),
                        "masks": lambda: ops.masks(q, b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid)}
            datax = ops.observe(q, b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid)[2]
            variants["cbf_filter(datax->h,J,u,r)"] = lambda: ops.cbf_filter(datax, goal, net, K, lo, hi, w["lam"], w["penalty"], 48, 2, torch.empty(0, device=dev))
            res["units"] = B
            res["unit"] = "states"
            for k, fn in variants.items():
                med, mn = timed(fn, args.iters, flush)
                res[k] = {"ms": med, "ms_min": mn, "per_s": B / (med * 1e-3)}
            if args.cpu and w["q"] is not None:
                rows = min(B, 1 << 16)
                cpu_fn = lambda: c_oracle.safety_filter(w["robot"], w["q"][:rows], w["boxes"], w["spheres"], True, w["weights"], 48, 2,
                                                        w["goal"], w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
                cpu_units = rows
        elif idx == 2:
            qf, qt = t(w["q"]), t(w["q_to"])
            E = qf.shape[0]
            fn = lambda: ops.edge_validity(qf, qt, w["n_interp"], b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid)
            med, mn = timed(fn, args.iters, flush)
            res.update(units=E, unit="edges", edge_validity={"ms": med, "ms_min": mn, "per_s": E / (med * 1e-3),
                                                              "configs_per_s": E * (w["n_interp"] + 1) / (med * 1e-3)})
            v = fn()[0]
            res["valid_fraction"] = float(v.float().mean().item())
            if args.cpu:
                rows = min(E, 1 << 12)
                cpu_fn = lambda: c_oracle.edge_validity(w["robot"], w["q"][:rows], w["q_to"][:rows], w["n_interp"], w["boxes"], w["spheres"], True)
                cpu_units = rows
        elif idx == 3:
            q0 = t(w["q"])
            T = q0.shape[0]
            safe = ops.masks(q0, b8, s4, 1, gr, w["thr_soft"], w["thr_hard"], rid)[0].bool()   # "filtered to safe" 
            keep = q0[safe]
            q0 = keep.repeat((T + keep.shape[0] - 1) // keep.shape[0], 1)[:T].contiguous()
            res["safe_start_fraction"] = float(safe.float().mean().item())
            fn = lambda: ops.rollout(q0, goal, w["steps"], w["ctrl_every"], w["dt"], 0, 10, 0.1, False, True, b8, s4, 1, gr, net, K, lo, hi,
                                     w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
            med, mn = timed(fn, max(2, args.iters // 2), flush)
            qf, alive, done, _ = fn()
            sim = float(done.sum().item())
            res.update(units=T, unit="trajectories",
                       rollout={"ms": med, "ms_min": mn, "traj_per_s": T / (med * 1e-3), "sim_steps_per_s": sim / (med * 1e-3),
                                "qp_steps_per_s": sim / w["ctrl_every"] / (med * 1e-3), "alive_fraction": float(alive.float().mean().item()),
                                "mean_steps": sim / T})
            if args.cpu:
                rows = min(T, 1 << 10)
                cpu_fn = lambda: c_oracle.rollout(w["robot"], w["q"][:rows], w["goal"][:rows], w["steps"], w["ctrl_every"], w["dt"], w["boxes"], None, True,
                                                  w["weights"], 48, 2, w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"], keep_going=True)
                cpu_units = rows
        if cpu_fn is not None:
            cpu_fn()
            t0 = time.perf_counter()
            cpu_fn()
            res["cpu_port"] = {"per_s": cpu_units / (time.perf_counter() - t0), "cores": threads, "sample_units": cpu_units}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
