#!/usr/bin/env python3
"""Device-resident timings of the planner front-end and the stand-alone QP entry points (1 GPU): cbfrrt::knn,
cbfrrt::radius_mask, cbfrrt::qp, cbfrrt::qp_backward, cbfrrt::qp_multi, cbfrrt::edge_validity_var.  One JSON line per measurement."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from cbfrrt import ops, workloads as W  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    out = []
    tree = t(rng.uniform(-2, 2, size=(100000, 7)))
    for M, k in ((50, 1), (1, 50), (4096, 1)):
        q = t(rng.uniform(-2, 2, size=(M, 7)))
        out.append(dict(op="knn", queries=M, tree=100000, k=k, ms=timed(lambda: ops.knn(q, tree, k))))
    q = t(rng.uniform(-2, 2, size=(200, 7)))
    out.append(dict(op="radius_mask", queries=200, points=100000, ms=timed(lambda: ops.radius_mask(q, tree, 1.0))))
    B, n = 1 << 20, 7
    lo, hi = t(-np.ones(n)), t(np.ones(n))
    a1, c1, ur = t(rng.normal(size=(B, n))), t(rng.normal(size=B)), t(rng.uniform(-1.5, 1.5, size=(B, n)))
    out.append(dict(op="qp", rows=B, n=n, ms=timed(lambda: ops.qp(a1, c1, ur, lo, hi, 5000.0))))
    gu, gr = t(rng.normal(size=(B, n))), t(rng.normal(size=B))
    ms = timed(lambda: ops.qp_backward(a1, c1, ur, lo, hi, 5000.0, gu, gr))
    out.append(dict(op="qp_backward", rows=B, n=n, ms=ms, algorithmic_bytes_per_row=4 * (3 * n + 2) + 4 * (2 * n + 1),
                    gb_per_s=B * (4 * (3 * n + 2) + 4 * (2 * n + 1)) / ms / 1e6))
    for S in (2, 4):
        a, c = t(rng.normal(size=(B, S, n))), t(rng.normal(scale=0.8, size=(B, S)))
        out.append(dict(op="qp_multi", rows=B, n=n, constraints=S, ms=timed(lambda: ops.qp_multi(a, c, ur, lo, hi, 5000.0, 100, 1e-9))))
        grS = t(rng.normal(size=(B, S)))
        out.append(dict(op="qp_multi_backward", rows=B, n=n, constraints=S,
                        ms=timed(lambda: ops.qp_multi_backward(a, c, ur, lo, hi, 5000.0, 100, 1e-9, gu, grS))))
    w = W.config(3, scale=1 / 4)
    rid = ops.ROBOT_ID[w["robot"]]
    b8, s4 = ops.pack_scene(w["boxes"][:, :3], w["boxes"][:, 3:], w["spheres"], device=dev)
    E = 1 << 18
    qf = t(W.sample_states("panda", E, 0))
    qt = t(np.clip(qf.cpu().numpy() + rng.uniform(-0.35, 0.35, size=(E, 7)).astype(np.float32), -2.8, 2.8))
    K = torch.from_numpy((np.linalg.norm((qf - qt).cpu().numpy(), axis=1) / 0.03).astype(np.int64))
    ms = timed(lambda: ops.edge_validity_var(qf, qt, K, b8, s4, 1, ops.no_grid(dev), 0.02, 0.0, rid), iters=5)
    out.append(dict(op="edge_validity_var", edges=E, points=int(K.sum() + E), boxes=int(b8.shape[0]), ms=ms))
    # data collection: EpisodicDataModule.sample_fixed and the noisy nominal rollouts on the batched ops
    import random
    import time
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.datamodules import EpisodicDataModule
    np.random.seed(5), random.seed(0), torch.manual_seed(0)
    env = ArmEnv(["panda"], config_file="", include_floor=True)
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.05, env=env, robot=env.robot_list[0])
    up, lo_ = dyn.state_limits
    dm = EpisodicDataModule(dyn, [(float(x), float(y)) for x, y in zip(lo_, up)], max_episode=1, trajectories_per_episode=16,
                            trajectory_length=400, fixed_samples=200000, noise_level=0.1,
                            quotas={"safe": 0.3, "unsafe": 0.3, "boundary": 0.2, "goal": 0.1})
    dm.sample_fixed()                                   # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, m = dm.sample_fixed()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out.append(dict(op="EpisodicDataModule.sample_fixed", robot="panda", boxes=int(env.obs_positions.shape[0]), rows=int(x.shape[0]),
                    wall_ms=dt * 1e3, labelled_rows_per_s=x.shape[0] / dt, note="host wall clock incl. rejection sampling and the 4 masks"))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    xs, ms = dm.sample_trajectories(dyn.noisy_simulator)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out.append(dict(op="EpisodicDataModule.sample_trajectories", trajectories=16, stored_states_per_trajectory=400, rows=int(xs.shape[0]),
                    wall_ms=dt * 1e3, stored_states_per_s=xs.shape[0] / dt, note="host wall clock; one observe launch per stored step"))
    # planner-sized calls: the reference's real regime is 1-50 states per call (eval_batchrrt_mindis_cbfsteer.py:66);
    # device time (CUDA events) and host wall time per call of the fused op, with the scene's broad phase built
    import time as _t
    w8 = W.config(3, scale=1e-9)
    b8p, s4p = ops.pack_scene(w8["boxes"][:, :3], w8["boxes"][:, 3:], w8["spheres"], device=dev)
    grp = ops.build_grid(b8p, s4p)
    netp = ops.pack_weights(w8["state_dict"], 7, 48, 2, device=dev)
    Kp, lop, hip, goalp = t(w8["K"]), t(w8["u_lo"]), t(w8["u_hi"]), t(w8["goal"][:1])
    for Bq in (1, 16, 64, 1024, 4096):
        qq = t(W.sample_states("panda", Bq, 3))
        for name, fn in (("safety_filter", lambda: ops.safety_filter(qq, goalp, b8p, s4p, 1, grp, netp, Kp, lop, hip, 0.1, 5000.0, 0.02, 0.0, 0, 48, 2)),
                         ("masks", lambda: ops.masks(qq, b8p, s4p, 1, grp, 0.02, 0.0, 0)),
                         ("safety_filter (no broad phase)", lambda: ops.safety_filter(qq, goalp, b8p, s4p, 1, ops.no_grid(dev), netp, Kp, lop, hip, 0.1, 5000.0, 0.02, 0.0, 0, 48, 2))):
            dev_ms = timed(fn, iters=20)
            torch.cuda.synchronize(); t0 = _t.perf_counter()
            for _ in range(200):
                fn()
            torch.cuda.synchronize(); wall = (_t.perf_counter() - t0) / 200
            out.append(dict(op=f"planner-sized {name}", robot="panda", boxes=8, states=Bq, device_us=dev_ms * 1e3, host_wall_us_per_call=wall * 1e6))
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
