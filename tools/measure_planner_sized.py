#!/usr/bin/env python3
"""Planner-sized calls (the reference's real regime: 1-50 states per call, eval_batchrrt_mindis_cbfsteer.py:66): device time
per call of every scene-consuming op with ONE STATE PER WARP (csrc/warp.cuh) against one thread per state, over the batch
size -- where the library's AUTO dispatch should switch.  Device time = CUDA-graph replay of 20 back-to-back calls between
two events (no host launch gaps); wall = eager calls from Python through ctypes.  One JSON line per (op, rows, path)."""
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

dev = "cuda"
t = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)


def device_us(fn, calls=20, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    class Eager:
        def replay(self):
            for _ in range(calls):
                fn()
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(calls):
                    fn()
    except Exception as e:   # not capturable: eager launches (includes host launch gaps)
        print(json.dumps({"note": f"graph capture failed, eager timing: {type(e).__name__}: {str(e)[:120]}"}), flush=True)
        torch.cuda.synchronize()
        g = Eager()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / calls)
    return best


def wall_us(fn, calls=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(calls):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / calls * 1e6


def main():
    sizes = [int(x) for x in os.environ.get("SIZES", "1,16,64,256,1024,4096,8192,16384").split(",")]
    for robot, cfg in (("panda", 3), ("magician", 1)):
        w = W.config(cfg, scale=1e-9)
        n, rid = w["n"], ops.ROBOT_ID[robot]
        boxes = W.scene(3)[0] if robot == "panda" else w["boxes"]
        sph = None if robot == "panda" else w["spheres"]
        b8, s4 = ops.pack_scene(boxes[:, :3], boxes[:, 3:], sph, device=dev)
        gr = ops.build_grid(b8, s4)
        net = ops.pack_weights(w["state_dict"], n, 48, 2, device=dev)
        K, lo, hi = t(w["K"]), t(w["u_lo"]), t(w["u_hi"])
        n_obs = int(b8.shape[0] + s4.shape[0])
        for B in sizes:
            q = t(W.sample_states(robot, B, 3))
            goal = t(W.sample_states(robot, 1, 4))
            E = max(B // 16, 1)
            qf, qt = t(W.sample_states(robot, E, 5)), t(W.sample_states(robot, E, 6))
            fns = {
                "safety_filter": lambda: ops.safety_filter(q, goal, b8, s4, 1, gr, net, K, lo, hi, w["lam"], w["penalty"], 0.02, 0.0, rid, 48, 2),
                "masks": lambda: ops.masks(q, b8, s4, 1, gr, 0.02, 0.0, rid),
                "observe": lambda: ops.observe(q, b8, s4, 1, gr, 0.02, 0.0, rid),
                "edge_validity(K=15)": lambda: ops.edge_validity(qf, qt, 15, b8, s4, 1, gr, 0.02, 0.0, rid),
            }
            for name, fn in fns.items():
                row = dict(op=name, robot=robot, obstacles=n_obs, rows=B if "edge" not in name else E * 16)
                for path in ("warp", "fma"):
                    ops.set_mlp_path(path)
                    row[f"{path}_device_us"] = round(device_us(fn), 2)
                    row[f"{path}_wall_us"] = round(wall_us(fn), 1)
                ops.set_mlp_path("auto")
                print(json.dumps(row), flush=True)
        # closed loop: T trajectories x 200 steps (BASELINE config 4's shape) and the planner's segmented expansion
        cand = W.sample_states(robot, 40000, 7)
        pool = cand[ops.masks(t(cand), b8, s4, 1, gr, 0.02, 0.0, rid)[0].bool().cpu().numpy()]
        for T in [int(x) for x in os.environ.get("TRAJ", "1,8,64,512,2048,4096,8192").split(",")]:
            q0 = t(np.resize(pool, (T, n)))
            goal = t(W.sample_states(robot, T, 8))
            fn = lambda: ops.rollout(q0, goal, 200, 4, 1 / 120, 9, 10, 3e-2, True, False, b8, s4, 1, gr, net, K, lo, hi, w["lam"],
                                     w["penalty"], 0.02, 0.0, rid, 48, 2)
            row = dict(op="rollout 200 steps", robot=robot, obstacles=n_obs, rows=T)
            for path in ("warp", "fma"):
                ops.set_mlp_path(path)
                row[f"{path}_device_us"] = round(device_us(fn, calls=3, reps=3), 1)
            ops.set_mlp_path("auto")
            print(json.dumps(row), flush=True)
        for T in (1, 10, 50, 128):
            q0 = t(np.resize(pool, (T, n)))
            goal = t(W.sample_states(robot, T, 9))
            fn = lambda: ops.rollout_segments(q0, goal, 9, 20, 4, 1 / 120, 9, 10, 3e-2, b8, s4, 1, gr, net, K, lo, hi, w["lam"], w["penalty"],
                                              0.02, 0.0, rid, 48, 2)
            row = dict(op="rollout_segments 9 x 20 steps", robot=robot, obstacles=n_obs, rows=T)
            for path in ("warp", "fma"):
                ops.set_mlp_path(path)
                row[f"{path}_device_us"] = round(device_us(fn, calls=5, reps=3), 1)
                row[f"{path}_wall_us"] = round(wall_us(fn, calls=30), 1)
            ops.set_mlp_path("auto")
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
