#!/usr/bin/env python3
"""Harness for the DRAFT tcgen05 PointNet kernel (pointnet_tc_draft.cu; never run on hardware in round 1).
Builds the draft as its own shared library (it is NOT part of libcbfrrt.so), runs it on random inputs on cuda:0 and
compares (1) the max-pooled 512-features and (2) the barrier value h with oracle/lidar_oracle.py (which is pinned against
the reference's own PointNet code, tests/test_lidar_oracle_cpu.py).  Bar: 1e-5 relative.

    /usr/local/graft/bin/gpurun -- python tools/experiments/pointnet_draft/pointnet_tc_draft_check.py
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
from oracle import lidar_oracle as L  # noqa: E402

N, S, R, PF, FD, HID = 7, 7, 128, 64, 32, 48


def rna_tf32(w):
    """cvt.rna.tf32.f32: round to nearest (ties away) onto 10 mantissa bits"""
    u = np.ascontiguousarray(w, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tiles(w):
    """W [n_out, K] (torch layout) -> [hi | lo] tiles in the no-swizzle K-major layout W4[k/4][n]"""
    w = np.ascontiguousarray(w, np.float32)
    hi = rna_tf32(w)
    lo = (w - hi).astype(np.float32)
    t = lambda m: m.reshape(m.shape[0], m.shape[1] // 4, 4).transpose(1, 0, 2).reshape(-1)
    return np.concatenate([t(hi), t(lo)])


def main():
    import torch
    so = os.path.join(HERE, "libpn_draft.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-diag-suppress", "20091", "-shared", "-o", so,
                           os.path.join(HERE, "pointnet_tc_draft.cu")])
    lib = C.CDLL(so)
    rng = np.random.default_rng(0)
    lin = lambda o, i: (rng.uniform(-1, 1, (o, i)).astype(np.float32) / np.sqrt(i).astype(np.float32),
                        rng.uniform(-1, 1, o).astype(np.float32) / np.sqrt(i).astype(np.float32))
    sd = {}
    for name, (o, i) in {"backbone_nn.fc11": (64, 13), "backbone_nn.fc2": (128, 64), "backbone_nn.fc3": (512, 128),
                         "backbone_fc.fc2": (256, 512), "backbone_fc.fc3": (PF, 256)}.items():
        sd[name + ".weight"], sd[name + ".bias"] = lin(o, i)
    enc = {}
    for name, (o, i) in {"encoder_nn.fc1": (512, S * PF), "encoder_nn.fc2": (256, 512), "encoder_nn.fc4": (FD, 256)}.items():
        enc[name + ".weight"], enc[name + ".bias"] = lin(o, i)
    hn = {}
    for name, (o, i) in {"input_linear": (HID, N + FD), "layer_0_linear": (HID, HID), "layer_1_linear": (HID, HID),
                         "output_linear": (1, HID)}.items():
        hn[name + ".weight"], hn[name + ".bias"] = lin(o, i)
    E = 16                                                     # state evaluations
    q = rng.uniform(-2, 2, (E, N)).astype(np.float32)
    pts = rng.uniform(-1, 1, (E, S, R, 3)).astype(np.float32)
    nrm = rng.normal(size=(E, S, R, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    obs = np.concatenate([pts, nrm], -1)                       # (E, S, R, 6), already in the sensor frames

    # ---- oracle
    x13 = np.concatenate([obs, np.broadcast_to(np.eye(S, dtype=np.float32)[None, :, None, :], (E, S, R, S))], -1)
    a = x13.reshape(E * S, R, 13)
    for nm in ("fc11", "fc2", "fc3"):
        a = L.leaky(L.linear(a, sd[f"backbone_nn.{nm}.weight"], sd[f"backbone_nn.{nm}.bias"]))
    max_o = a.max(axis=1)                                      # (E*S, 512)
    feat = L.pointnet_feat(obs.reshape(E, -1), sd, S, R)
    h_o = L.h_nn(np.concatenate([q, L.encoder(feat, enc)], 1), hn, 2)

    # ---- device
    dev = torch.device("cuda", 0)
    t = lambda v: torch.as_tensor(np.ascontiguousarray(v), device=dev)
    x16 = np.zeros((E * S * R, 16), np.float32)
    x16[:, :13] = x13.reshape(-1, 13)
    w1 = np.zeros((64, 16), np.float32)
    w1[:, :13] = sd["backbone_nn.fc11.weight"]
    w3 = sd["backbone_nn.fc3.weight"]
    w3t = np.concatenate([tiles(w3[c * 64:(c + 1) * 64]) for c in range(8)])
    bias = np.concatenate([sd["backbone_nn.fc11.bias"], sd["backbone_nn.fc2.bias"], sd["backbone_nn.fc3.bias"]])
    d_x, d_w1, d_w2, d_w3, d_b = t(x16), t(tiles(w1)), t(tiles(sd["backbone_nn.fc2.weight"])), t(w3t), t(bias)
    d_max = torch.full((E * S, 512), -2 ** 31, dtype=torch.int32, device=dev)
    P = lambda v: C.c_void_p(v.data_ptr())
    lib.pn_backbone.argtypes = [C.c_void_p, C.c_longlong, C.c_int] + [C.c_void_p] * 6
    rc = lib.pn_backbone(P(d_x), E * S * R, R, P(d_w1), P(d_w2), P(d_w3), P(d_b), P(d_max), None)
    torch.cuda.synchronize()
    assert rc == 0, rc
    keys = d_max.cpu().numpy()
    got = np.where(keys >= 0, keys, keys ^ 0x7FFFFFFF).astype(np.int32).view(np.float32)
    err = np.abs(got - max_o) / np.maximum(np.abs(max_o), 0.1)
    print("max-pooled features: max rel err", err.max())

    class HeadW(C.Structure):
        _fields_ = [(k, C.c_void_p) for k in ("fc2_w", "fc2_b", "fc3_w", "fc3_b", "e1_w", "e1_b", "e2_w", "e2_b", "e4_w", "e4_b",
                                              "h0_w", "h0_b", "h1_w", "h1_b", "h2_w", "h2_b", "ho_w", "ho_b")]
    keep = [t(v) for v in (sd["backbone_fc.fc2.weight"], sd["backbone_fc.fc2.bias"], sd["backbone_fc.fc3.weight"],
                           sd["backbone_fc.fc3.bias"], enc["encoder_nn.fc1.weight"], enc["encoder_nn.fc1.bias"],
                           enc["encoder_nn.fc2.weight"], enc["encoder_nn.fc2.bias"], enc["encoder_nn.fc4.weight"],
                           enc["encoder_nn.fc4.bias"], hn["input_linear.weight"], hn["input_linear.bias"],
                           hn["layer_0_linear.weight"], hn["layer_0_linear.bias"], hn["layer_1_linear.weight"],
                           hn["layer_1_linear.bias"], hn["output_linear.weight"], hn["output_linear.bias"])]
    hw = HeadW(*[v.data_ptr() for v in keep])
    d_q, d_h = t(q), torch.empty(E, dtype=torch.float32, device=dev)
    lib.pn_head.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.pn_head(P(d_max), P(d_q), E, N, S, PF, FD, C.byref(hw), P(d_h), None)
    torch.cuda.synchronize()
    assert rc == 0, rc
    eh = np.abs(d_h.cpu().numpy() - h_o[:, 0]) / np.maximum(np.abs(h_o[:, 0]), 0.1)
    print("h: max rel err", eh.max())
    assert err.max() < 1e-5 and eh.max() < 1e-5


if __name__ == "__main__":
    main()
