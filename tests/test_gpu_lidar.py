"""GPU suite (-m gpu): the LiDAR / point-cloud barrier (cbfrrt_lidar_cbf, csrc/lidar.cuh: tcgen05
PointNet backbone + FP32 head) against (1) the vectors recorded from the reference's own ArmLidar / PointNetfeat /
PointNetVanillaEncoder / NeuralLidarCBFController (tests/golden/ref_lidar_panda.npz, made by tests/golden/make_lidar_golden.py)
and (2) the numpy oracle (oracle/lidar_oracle.py, itself pinned to those vectors in the CPU suite) on larger random batches."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_lidar_panda.npz"), allow_pickle=False)
CFG = json.loads(str(G["cfg"]))
N = S = 7
TOL = 1e-5


def rel(a, b, floor=0.1):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b) / np.maximum(np.abs(b), floor)))


def _controller():
    from cbfrrt.neural_cbf.controllers import NeuralLidarCBFController, LidarLayout
    dm = LidarLayout(n_dims=N, list_sensor=list(range(S)), ray_per_sensor=CFG["n_obs"], add_normal=CFG["add_normal"],
                     controller_dt=float(G["controller_dt"]))
    torch.manual_seed(CFG["weight_seed"])
    ctrl = NeuralLidarCBFController(dm, **{k: CFG[k] for k in ("per_feature_dim", "feature_dim", "cbf_hidden_layers",
                                                                "cbf_hidden_size", "use_bn", "use_neural_actor")})
    sd = ctrl.state_dict()
    count = sum(v.numel() for v in sd.values())
    total = float(sum(v.double().abs().sum() for v in sd.values()))
    if count != int(G["param_count"]) or abs(total - float(G["param_abs_sum"])) > 1e-6 * float(G["param_abs_sum"]):
        pytest.skip("torch's nn.Linear initialisation differs from the one the golden vectors were recorded with")
    return ctrl


def test_lidar_h_and_jacobian_match_reference_golden():
    """h and the one-sided numerical Jacobian of the reference's own NeuralLidarCBFController (R = 64 rays x 7 sensors)"""
    ctrl = _controller()
    datax = torch.as_tensor(G["datax"])
    h = ctrl.h(datax, sampled_index=G["sampled_index"])
    assert h.shape == (6, 1) and rel(h.cpu().numpy(), G["h"]) < TOL
    hj, J, _ = ctrl.h_with_jacobian(datax, (torch.as_tensor(G["J_p"]), torch.as_tensor(G["J_R"])), sampled_index=G["sampled_index_jac"])
    assert rel(hj.cpu().numpy(), G["h_jac"]) < TOL
    # J = (h' - h) / (2 d), d = 1/60: a difference of two float32 evaluations divided by 1/30 -- absolute 1e-5 * 30
    assert J.shape == (6, 1, N) and np.abs(J.cpu().numpy() - G["J"]).max() < 3e-4
    # the reference's own draw of the ray indices, replayed through the same generator calls
    torch.manual_seed(CFG["index_seed"])
    assert rel(ctrl.h(datax).cpu().numpy(), G["h"]) < TOL


@pytest.mark.parametrize("B,R,P", [(37, 128, 500), (64, 64, 300), (5, 1024, 4096), (200, 32, 64)])
def test_lidar_matches_oracle_random_batches(B, R, P):
    """larger batches, other ray counts (tiles that straddle evaluations, a partly idle tail tile), against the oracle"""
    from oracle import lidar_oracle as L
    ctrl = _controller()
    ctrl.dynamics_model.ray_per_sensor = R
    rng = np.random.default_rng(B + R)
    q = rng.uniform(-2, 2, (B, N)).astype(np.float32)
    pts = rng.uniform(-1, 1, (B, P, 3)).astype(np.float32)
    nrm = rng.normal(size=(B, P, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    rot = np.linalg.qr(rng.normal(size=(B, S, 3, 3)))[0].astype(np.float32)
    org = (rng.uniform(size=(B, S, 3)) - 0.5).astype(np.float32)
    datax = np.concatenate([q, np.concatenate([pts, nrm], -1).reshape(B, -1), np.concatenate([org, rot.reshape(B, S, 9)], -1).reshape(B, -1)], 1)
    idx = rng.integers(0, P, (S, R)).astype(np.int32)
    J_p = (0.3 * rng.normal(size=(B, S, 3, N))).astype(np.float32)
    J_R = (0.3 * rng.normal(size=(B, S, 3, 3, N))).astype(np.float32)
    np_sd = lambda mod: {k: v.detach().numpy() for k, v in mod.state_dict().items()}
    nets = {"pc_head": np_sd(ctrl.pc_head), "encoder": np_sd(ctrl.encoder), "h_nn": np_sd(ctrl.h_nn)}
    h_o = L.h(datax, nets, N, S, R, idx, layers=CFG["cbf_hidden_layers"])
    h = ctrl.h(torch.as_tensor(datax), sampled_index=idx)
    assert rel(h.cpu().numpy(), h_o) < TOL
    if B <= 64 and R <= 128:
        ho2, J_o = L.h_with_jacobian(datax, J_p, J_R, nets, N, S, R, idx, float(G["controller_dt"]), layers=CFG["cbf_hidden_layers"])
        hj, J, _ = ctrl.h_with_jacobian(torch.as_tensor(datax), (torch.as_tensor(J_p), torch.as_tensor(J_R)), sampled_index=idx)
        assert rel(hj.cpu().numpy(), ho2) < TOL and np.abs(J.cpu().numpy() - J_o).max() < 3e-4


def test_lidar_argument_validation():
    from cbfrrt import ops
    from cbfrrt._lib import CbfrrtError
    ctrl = _controller()
    datax = torch.as_tensor(G["datax"])
    with pytest.raises(CbfrrtError):       # rays must be a multiple of 32
        ctrl.h(datax, sampled_index=G["sampled_index"][:, :48])
    with pytest.raises(NotImplementedError):
        ops.lidar_cbf(ctrl._net(), torch.zeros(2, 7), torch.zeros(2, 10, 6), torch.zeros(2, 7, 12), torch.zeros(7, 64, dtype=torch.int32))
