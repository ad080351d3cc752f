"""LiDAR / point-cloud barrier path: the oracle (the kernels are checked in tests/test_gpu_lidar.py).
oracle/lidar_oracle.py (numpy) against the golden vectors recorded from the reference's own ArmLidar / PointNetfeat /
PointNetVanillaEncoder / NeuralLidarCBFController (tests/golden/make_lidar_golden.py -> ref_lidar_panda.npz)."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import lidar_oracle as L

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_lidar_panda.npz"), allow_pickle=False)
CFG = json.loads(str(G["cfg"]))
N = S = 7
R = CFG["n_obs"]
TOL = 1e-5


def rel(a, b, floor=1.0):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b) / np.maximum(np.abs(b), floor)))


@pytest.fixture(scope="module")
def nets():
    """the reference's modules rebuilt from the recorded seed in the reference's construction order
    (neural_lidar_cbf_controller.py:66-95, pointnet.py:20-48,78-90); a checksum guards against RNG drift"""
    torch.manual_seed(CFG["weight_seed"])
    cin, pf = 3 + 3 * int(CFG["add_normal"]) + S, CFG["per_feature_dim"]
    backbone_nn = nn.Sequential(OrderedDict([("fc11", nn.Linear(cin, 64)), ("fc2", nn.Linear(64, 128)), ("fc3", nn.Linear(128, 512))]))
    backbone_fc = nn.Sequential(OrderedDict([("fc2", nn.Linear(512, 256)), ("fc3", nn.Linear(256, pf))]))
    encoder_nn = nn.Sequential(OrderedDict([("fc1", nn.Linear(pf * S, 512)), ("fc2", nn.Linear(512, 256)),
                                            ("fc4", nn.Linear(256, CFG["feature_dim"]))]))
    H = CFG["cbf_hidden_size"]
    h_layers = OrderedDict([("input_linear", nn.Linear(N + CFG["feature_dim"], H))])
    for i in range(CFG["cbf_hidden_layers"]):
        h_layers[f"layer_{i}_linear"] = nn.Linear(H, H)
    h_layers["output_linear"] = nn.Linear(H, 1)
    h_nn = nn.Sequential(h_layers)
    np_sd = lambda mod, prefix: {f"{prefix}{k}": v.detach().numpy() for k, v in mod.state_dict().items()}
    nets = {"pc_head": {**np_sd(backbone_nn, "backbone_nn."), **np_sd(backbone_fc, "backbone_fc.")},
            "encoder": np_sd(encoder_nn, "encoder_nn."), "h_nn": np_sd(h_nn, "")}
    count = sum(v.size for d in nets.values() for v in d.values())
    total = float(sum(np.abs(v.astype(np.float64)).sum() for d in nets.values() for v in d.values()))
    if count != int(G["param_count"]) or abs(total - float(G["param_abs_sum"])) > 1e-6 * float(G["param_abs_sum"]):
        pytest.skip("torch's nn.Linear initialisation differs from the one the golden vectors were recorded with")
    return nets


def test_parameter_inventory(nets):
    assert int(G["param_count"]) == 599153       # 13-64-128-512 | 512-256-64 | 448-512-256-32 | 39-48-48-48-1
    assert sorted(G["param_names"].tolist())[0].startswith("encoder.encoder_nn")


def test_datax_to_x_matches_reference():
    x = L.datax_to_x(G["datax"], N, S, R, G["sampled_index"], add_normal=True)
    assert x.shape == G["x"].shape == (6, N + S * R * 6)
    assert rel(x, G["x"]) < TOL


def test_pointnet_and_encoder_match_reference(nets):
    feat = L.pointnet_feat(G["x"][:, N:], nets["pc_head"], S, R)
    assert feat.shape == (6, S * CFG["per_feature_dim"]) and rel(feat, G["feature"], floor=0.1) < TOL
    enc = L.encoder(G["feature"], nets["encoder"])
    assert enc.shape == (6, CFG["feature_dim"]) and rel(enc, G["encoded"], floor=0.1) < TOL


def test_h_matches_reference(nets):
    h = L.h(G["datax"], nets, N, S, R, G["sampled_index"], layers=CFG["cbf_hidden_layers"])
    assert h.shape == (6, 1) and rel(h, G["h"], floor=0.1) < TOL


def test_lookahead_and_numerical_jacobian_match_reference(nets):
    nxt = L.lookahead(G["datax"], G["dq"], G["J_p"], G["J_R"], N, S)
    assert rel(nxt, G["lookahead"]) < TOL
    h, J = L.h_with_jacobian(G["datax"], G["J_p"], G["J_R"], nets, N, S, R, G["sampled_index_jac"], float(G["controller_dt"]),
                             layers=CFG["cbf_hidden_layers"])
    assert rel(h, G["h_jac"], floor=0.1) < TOL
    # J = (h' - h) / (2 d), d = 1/60: a difference of two float32 evaluations divided by 1/30 -- absolute 1e-5 * 30
    assert J.shape == (6, 1, N) and np.abs(J - G["J"]).max() < 3e-4
