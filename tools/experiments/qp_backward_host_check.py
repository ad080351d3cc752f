#!/usr/bin/env python3
"""Runs the HOST build of the device reverse-mode QP (qp_backward_host.cu) against the oracle adjoint on random rows.
    nvcc -O2 -o tools/experiments/qp_backward_host tools/experiments/qp_backward_host.cu && python tools/experiments/qp_backward_host_check.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle  # noqa: E402

exe = os.path.join(ROOT, "tools", "experiments", "qp_backward_host")
worst = 0.0
for n, p in [(7, 5000.0), (4, 5000.0), (7, 0.5), (1, 2.0), (2, 50.0), (3, 1e3), (5, 0.05), (6, 1e9), (8, 5000.0), (7, 1e9)]:
    rng = np.random.default_rng(200 + n)
    B = 200000
    f = lambda x: np.ascontiguousarray(x, np.float32)
    lo, hi = f(-rng.uniform(0.5, 2.0, n)), f(rng.uniform(0.5, 2.0, n))
    a, c, ur = f(rng.normal(size=(B, n))), f(rng.normal(scale=2.0, size=B)), f(rng.uniform(-2.5, 2.5, size=(B, n)))
    gu, gr = f(rng.normal(size=(B, n))), f(rng.normal(size=B))
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "in"), "wb") as fh:
            for x in (a, c, ur, gu, gr, lo, hi):
                fh.write(x.tobytes())
        subprocess.check_call([exe, str(n), str(B), repr(p), os.path.join(d, "in"), os.path.join(d, "out")])
        out = np.fromfile(os.path.join(d, "out"), np.float32)
    ga, gc, gt = out[:B * n].reshape(B, n), out[B * n:B * n + B], out[B * n + B:].reshape(B, n)
    ga_o, gc_o, gt_o, margin = c_oracle.qp_backward(a, c, ur, lo, hi, p, gu, gr, want_margin=True)
    scale = np.maximum(1.0, np.maximum(np.abs(ga_o).max(1), np.maximum(np.abs(gt_o).max(1), np.abs(gc_o))))
    err = np.maximum(np.abs(ga - ga_o).max(1), np.maximum(np.abs(gt - gt_o).max(1), np.abs(gc - gc_o))) / scale
    ok = margin > 1e-6
    print(f"n={n} p={p:g}: rows off a kink {ok.mean():.5f}, max err there {err[ok].max():.2e}, rows differing at kinks {(err[~ok] > 1e-5).sum()}")
    worst = max(worst, err[ok].max())
assert worst < 1e-5, worst
print("ok")
