#!/usr/bin/env python3
"""HOST build of the device reverse mode of the multi-scenario QP (qp_multi_backward_host.cu) against the oracle adjoint.
    nvcc -O2 -w -o tools/experiments/qp_multi_backward_host tools/experiments/qp_multi_backward_host.cu
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle  # noqa: E402

exe = os.path.join(ROOT, "tools", "experiments", "qp_multi_backward_host")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
for n, S, p in [(7, 2, 50.0), (7, 3, 1e3), (4, 2, 0.5), (7, 4, 0.05), (7, 2, 1e9), (1, 2, 50.0), (2, 4, 5000.0), (8, 3, 1e3), (7, 1, 50.0), (5, 4, 50.0)]:
    rng = np.random.default_rng(300 + n * S)
    f = lambda x: np.ascontiguousarray(x, np.float32)
    lo, hi = f(-rng.uniform(0.5, 2.0, n)), f(rng.uniform(0.5, 2.0, n))
    a, c, ur = f(rng.normal(size=(B, S, n))), f(rng.normal(scale=0.8, size=(B, S))), f(rng.uniform(-2.0, 2.0, size=(B, n)))
    gu, gr = f(rng.normal(size=(B, n))), f(rng.normal(size=(B, S)))
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "in"), "wb") as fh:
            for x in (a, c, ur, gu, gr, lo, hi):
                fh.write(x.tobytes())
        subprocess.check_call([exe, str(n), str(S), str(B), repr(p), os.path.join(d, "in"), os.path.join(d, "out")])
        out = np.fromfile(os.path.join(d, "out"), np.float32)
    ga, gc, gt = out[:B * S * n].reshape(B, S, n), out[B * S * n:B * S * n + B * S].reshape(B, S), out[B * S * n + B * S:].reshape(B, n)
    ga_o, gc_o, gt_o, margin = c_oracle.qp_multi_backward(a, c, ur, lo, hi, p, gu, gr, want_margin=True)
    scale = np.maximum(1.0, np.maximum(np.abs(ga_o).max((1, 2)), np.maximum(np.abs(gt_o).max(1), np.abs(gc_o).max(1))))
    err = np.maximum(np.abs(ga - ga_o).max((1, 2)), np.maximum(np.abs(gt - gt_o).max(1), np.abs(gc - gc_o).max(1))) / scale
    ok = margin > 1e-5
    print(f"n={n} S={S} p={p:g}: rows off a kink {ok.mean():.5f}, max err there {err[ok].max():.2e}, "
          f"rows > 1e-5 there {(err[ok] > 1e-5).sum()}, rows differing near kinks {(err[~ok] > 1e-5).sum()}")
