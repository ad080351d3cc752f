#!/usr/bin/env python3
"""Device-resident timing of the LiDAR / PointNet barrier (cbfrrt_lidar_cbf): states/s, points/s and the tensor-pipe
fraction of the backbone against a dense TF32 matmul measured in the same run.
    python tools/measure_lidar.py [--states 256] [--rays 1024] [--cloud 8192] [--no-jacobian]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from cbfrrt.neural_cbf.controllers import NeuralLidarCBFController, LidarLayout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=256)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--cloud", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-jacobian", action="store_true")
    a = ap.parse_args()
    n = S = 7
    dev = torch.device("cuda", 0)
    dm = LidarLayout(n_dims=n, list_sensor=list(range(S)), ray_per_sensor=a.rays)
    torch.manual_seed(0)
    ctrl = NeuralLidarCBFController(dm, per_feature_dim=64, feature_dim=32, cbf_hidden_layers=2, cbf_hidden_size=48, use_bn=False)
    B, P = a.states, a.cloud
    g = torch.Generator().manual_seed(1)
    q = torch.rand(B, n, generator=g) * 4 - 2
    cloud = torch.cat([torch.rand(B, P, 3, generator=g) * 2 - 1, torch.nn.functional.normalize(torch.randn(B, P, 3, generator=g), dim=-1)], -1)
    rot, _ = torch.linalg.qr(torch.randn(B, S, 3, 3, generator=g))
    frames = torch.cat([torch.rand(B, S, 3, generator=g) - 0.5, rot.reshape(B, S, 9)], -1)
    datax = torch.cat([q, cloud.reshape(B, -1), frames.reshape(B, -1)], 1).to(dev)
    jac = (0.3 * torch.randn(B, S, 3, n, generator=g).to(dev), 0.3 * torch.randn(B, S, 3, 3, n, generator=g).to(dev))
    idx = torch.randint(0, P, (S, a.rays), generator=g).int().to(dev)
    fn = (lambda: ctrl.h(datax, sampled_index=idx)) if a.no_jacobian else (lambda: ctrl.h_with_jacobian(datax, jac, sampled_index=idx))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ms = float(np.median([s.elapsed_time(e) for s, e in ev]))
    evals = B * (1 if a.no_jacobian else n + 1)
    points = evals * S * a.rays
    flop_pt = 2 * (13 * 64 + 64 * 128 + 128 * 512)            # algorithmic, per point (pointnet.py:24-26)
    exec_pt = 3 * 2 * (16 * 64 + 64 * 128 + 128 * 512)        # executed on the tensor cores: K padded to 16, 3-term TF32 split
    torch.backends.cuda.matmul.allow_tf32 = True
    x = torch.randn(8192, 8192, device=dev); y = torch.randn(8192, 8192, device=dev)
    (x @ y); tt = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        e0.record(); (x @ y); e1.record(); torch.cuda.synchronize(); tt.append(e0.elapsed_time(e1))
    tf32_peak = 2 * 8192 ** 3 / (min(tt) * 1e-3) / 1e12
    print(json.dumps({"states": B, "rays_per_sensor": a.rays, "sensors": S, "evaluations_per_state": evals // B, "ms": ms,
                      "states_per_s": B / (ms * 1e-3), "points_per_s": points / (ms * 1e-3),
                      "algorithmic_tflops": flop_pt * points / (ms * 1e-3) / 1e12,
                      "executed_tensor_tflops": exec_pt * points / (ms * 1e-3) / 1e12, "dense_tf32_tflops_measured": tf32_peak,
                      "tensor_pipe_frac": exec_pt * points / (ms * 1e-3) / 1e12 / tf32_peak,
                      "note": "whole call (preset + backbone + head + finish); the backbone kernel dominates"}))


if __name__ == "__main__":
    main()
