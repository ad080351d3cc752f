#!/usr/bin/env python3
"""Golden vectors of the LiDAR / point-cloud barrier path, recorded by EXECUTING THE REFERENCE'S OWN
PYTHON in the dev container (same inert stubs as make_golden.py for pybullet / Lightning / cvxpylayers):

    neural_cbf/systems/arm_lidar.py:219-248           ArmLidar.datax_to_x
    neural_cbf/systems/arm_lidar.py:250-288           ArmLidar.batch_lookahead / _jacobian_to_auxlookahead
    neural_cbf/controllers/utils/pointnet.py:13-97    PointNetfeat, PointNetVanillaEncoder
    neural_cbf/controllers/neural_lidar_cbf_controller.py:115-219   h, h_with_jacobian

Inputs are synthetic (a random world-frame cloud with normals, random orthonormal sensor frames, random frame
Jacobians): the vectors pin the reference's ARITHMETIC on given inputs, not pybullet's FK or surface sampling.
The reference samples the per-sensor point indices with torch.randint inside datax_to_x; they are recorded by replaying
the generator state.  Weights are regenerated from the seed by the test (a checksum is stored), so the fixture stays small.

Output: tests/golden/ref_lidar_panda.npz.   Run: python tests/golden/make_lidar_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

CFG = dict(n_obs=64, point_in_dataset_pc=128, point_dim=3, add_normal=True, per_feature_dim=64, feature_dim=32,
           cbf_hidden_layers=2, cbf_hidden_size=48, use_bn=False, use_neural_actor=False, cbf_alpha=0.1,
           cbf_relaxation_penalty=5000.0, safe_level=0.1, unsafe_level=0.1, weight_seed=11, input_seed=12, index_seed=13)


def build_reference(m):
    from neural_cbf.systems import ArmLidar
    from neural_cbf.controllers.neural_lidar_cbf_controller import NeuralLidarCBFController
    env, robot = mg._FakeEnv(m), mg._FakeRobot(m)
    dyn = ArmLidar({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=robot,
                   n_obs=CFG["n_obs"], point_in_dataset_pc=CFG["point_in_dataset_pc"], list_sensor=robot.body_joints,
                   observation_type="uniform_surface", point_dim=CFG["point_dim"], add_normal=CFG["add_normal"])
    torch.manual_seed(CFG["weight_seed"])
    ctrl = NeuralLidarCBFController(dyn, [{"m1": 5.76}], None, None,
                                    **{k: CFG[k] for k in ("per_feature_dim", "feature_dim", "cbf_hidden_layers",
                                                           "cbf_hidden_size", "use_bn", "use_neural_actor", "cbf_alpha",
                                                           "cbf_relaxation_penalty", "safe_level", "unsafe_level")})
    ctrl.eval()
    return dyn, ctrl


def make_inputs(m, B=6):
    n, S, P = m["n"], m["n"], CFG["point_in_dataset_pc"]
    g = torch.Generator().manual_seed(CFG["input_seed"])
    qlo, qhi = torch.tensor(m["q_lower"]).float(), torch.tensor(m["q_upper"]).float()
    q = qlo + (qhi - qlo) * torch.rand(B, n, generator=g)
    pts = torch.rand(B, P, 3, generator=g) * 2 - 1
    nrm = torch.nn.functional.normalize(torch.randn(B, P, 3, generator=g), dim=-1)
    cloud = torch.cat([pts, nrm], dim=-1)
    rot, _ = torch.linalg.qr(torch.randn(B, S, 3, 3, generator=g))
    org = torch.rand(B, S, 3, generator=g) - 0.5
    aux = torch.cat([org, rot.reshape(B, S, 9)], dim=-1)
    datax = torch.cat([q, cloud.reshape(B, -1), aux.reshape(B, -1)], dim=1)
    J_p = 0.3 * torch.randn(B, S, 3, n, generator=g)
    J_R = 0.3 * torch.randn(B, S, 3, 3, n, generator=g)
    return datax, J_p, J_R


def main():
    sys.meta_path.insert(0, mg._StubFinder())
    sys.path.insert(0, mg.REF)
    m = json.load(open(os.path.join(mg.ROOT, "spec", "robot_models.json")))["panda"]
    dyn, ctrl = build_reference(m)
    n, S, R, P = m["n"], m["n"], CFG["n_obs"], CFG["point_in_dataset_pc"]
    datax, J_p, J_R = make_inputs(m)
    out = {}

    def indices(seed):                        # what datax_to_x will draw (arm_lidar.py:236-238), replayed
        torch.manual_seed(seed)
        return np.stack([torch.randint(low=0, high=P, size=(R, 1)).squeeze().int().numpy() for _ in range(S)])

    idx = indices(CFG["index_seed"])
    torch.manual_seed(CFG["index_seed"])
    x = dyn.datax_to_x(datax)
    out.update(sampled_index=idx, x=x.numpy())
    with torch.no_grad():
        feat = ctrl.pc_head(x[:, n:])
        enc = ctrl.encoder(feat)
        out.update(feature=feat.numpy(), encoded=enc.numpy())
        torch.manual_seed(CFG["index_seed"])
        out["h"] = ctrl.h(datax).numpy()
        dq = 0.01 * torch.randn(datax.shape[0], n, generator=torch.Generator().manual_seed(5))
        out.update(dq=dq.numpy(), lookahead=dyn.batch_lookahead(datax, dq, (J_p, J_R)).numpy())
    torch.manual_seed(CFG["index_seed"] + 1)
    hj, J, _ = ctrl.h_with_jacobian(datax, (J_p, J_R))
    out.update(sampled_index_jac=indices(CFG["index_seed"] + 1), h_jac=hj.detach().numpy(), J=J.detach().numpy())
    sd = {k: v.detach().numpy() for k, v in ctrl.state_dict().items()}
    checksum = float(sum(np.abs(v.astype(np.float64)).sum() for v in sd.values()))
    np.savez_compressed(os.path.join(HERE, "ref_lidar_panda.npz"), datax=datax.numpy(), J_p=J_p.numpy(), J_R=J_R.numpy(),
                        cfg=json.dumps(CFG), param_names=np.array(sorted(sd)), param_count=sum(v.size for v in sd.values()),
                        param_abs_sum=checksum, controller_dt=dyn.controller_dt, **out)
    print("params", sum(v.size for v in sd.values()), "abs-sum", checksum, "h", out["h"].ravel()[:3], "J", out["J"][0, 0, :3])


if __name__ == "__main__":
    main()
