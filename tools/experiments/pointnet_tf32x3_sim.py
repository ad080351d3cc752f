#!/usr/bin/env python3
"""Design check for the planned tcgen05 PointNet kernel, CPU only: is the 3-term TF32 split
(hi*hi + hi*lo + lo*hi, operands truncated to 10 mantissa bits as kind::tf32 does, fp32 accumulation) accurate enough
for the 1e-5 parity bar on the reference's point-cloud barrier?  Emulates the split in numpy on the golden inputs of
tests/golden/ref_lidar_panda.npz and compares h with the reference's recorded float32 value; also plain TF32 (1 term)
and bf16x3 for contrast.     python tools/experiments/pointnet_tf32x3_sim.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import lidar_oracle as L  # noqa: E402


def trunc(x, keep_bits):
    """keep the sign, exponent and the top ``keep_bits`` mantissa bits of float32 (tf32: 10, bf16: 7)"""
    u = np.asarray(x, np.float32).view(np.uint32)
    return (u & np.uint32((0xFFFFFFFF << (23 - keep_bits)) & 0xFFFFFFFF)).view(np.float32)


def make_linear(terms, keep_bits):
    def linear(x, w, b):
        x, w = np.asarray(x, np.float32), np.asarray(w, np.float32)
        xh, wh = trunc(x, keep_bits), trunc(w, keep_bits)
        acc = (xh.astype(np.float64) @ wh.astype(np.float64).T)
        if terms >= 3:
            xl, wl = trunc(x - xh, keep_bits), trunc(w - wh, keep_bits)
            acc = acc + xh.astype(np.float64) @ wl.astype(np.float64).T + xl.astype(np.float64) @ wh.astype(np.float64).T
        return (acc.astype(np.float32) + np.asarray(b, np.float32)).astype(np.float32)   # fp32 accumulator read-back + bias
    return linear


def main():
    import torch  # weights are rebuilt like tests/test_lidar_oracle_cpu.py does
    import test_lidar_oracle_cpu as T
    nets = T.nets.__wrapped__() if hasattr(T.nets, "__wrapped__") else None
    if nets is None:
        raise SystemExit("could not rebuild the weights")
    G, CFG = T.G, T.CFG
    ref = G["h"].astype(np.float64)
    exact = L.linear
    rows = []
    for name, lin in (("fp32 (oracle)", exact), ("tf32 x1", make_linear(1, 10)), ("tf32 x3", make_linear(3, 10)),
                      ("bf16 x3", make_linear(3, 7))):
        L.linear = lin
        try:
            h = L.h(G["datax"], nets, 7, 7, CFG["n_obs"], G["sampled_index"], layers=CFG["cbf_hidden_layers"])
            feat = L.pointnet_feat(G["x"][:, 7:], nets["pc_head"], 7, CFG["n_obs"])
        finally:
            L.linear = exact
        rows.append((name, float(np.abs(h - ref).max() / np.maximum(np.abs(ref), 0.1).max()),
                     float(np.max(np.abs(feat - G["feature"]) / np.maximum(np.abs(G["feature"]), 0.1)))))
    for name, eh, ef in rows:
        print(f"{name:14s} max rel err of h {eh:.2e}   of the PointNet feature {ef:.2e}")
    print(json.dumps({r[0]: {"h": r[1], "feature": r[2]} for r in rows}))


if __name__ == "__main__":
    main()
