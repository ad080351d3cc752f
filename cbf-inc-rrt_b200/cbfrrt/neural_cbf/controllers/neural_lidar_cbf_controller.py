"""LiDAR / point-cloud barrier -- mirror of neural_cbf/controllers/neural_lidar_cbf_controller.py:27-219 (inference
methods ``h`` and ``h_with_jacobian``) over cbfrrt::lidar_cbf (include/cbfrrt.h cbfrrt_lidar_cbf).

The reference evaluates, per call, ArmLidar.datax_to_x (arm_lidar.py:219-248), PointNetfeat / PointNetVanillaEncoder
(controllers/utils/pointnet.py:13-97), h_nn, and for the Jacobian 2 n look-ahead copies of every row
(batch_lookahead, arm_lidar.py:250-288) of which it uses n (:205-213).  Here one library call does all of it for the
whole batch; the look-ahead rows are never materialised.

What stays outside (out of scope): producing the point clouds (pybullet ray casting / surface sampling,
arm_lidar.py:92-205) and the training losses.  ``dynamics_model`` therefore only has to describe the data layout:
``n_dims``, ``list_sensor``, ``ray_per_sensor``, ``point_dims``, ``add_normal``, ``controller_dt`` -- the attributes the
reference's own controller reads (LidarLayout below is enough; the reference's ArmLidar has them too).

datax layout (arm_lidar.py:219-226): [q (n) | cloud (P x point_dims, world frame) | per-sensor frames (S x (origin 3 |
rotation 9))].  The reference draws the sampled ray indices with torch.randint inside datax_to_x; ``h`` draws them the
same way (same generator calls, so a seeded call reproduces the reference's draw) unless ``sampled_index`` is given.
"""
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from ... import ops


@dataclass
class LidarLayout:
    """the attributes of the reference's ArmLidar that the controller reads"""
    n_dims: int
    list_sensor: List[int]
    ray_per_sensor: int
    add_normal: bool = True
    controller_dt: float = 1.0 / 30.0
    point_dim: int = 3
    device: str = "cuda"
    q_dims: int = field(init=False)
    point_dims: int = field(init=False)

    def __post_init__(self):
        self.q_dims = self.n_dims
        self.point_dims = self.point_dim + 3 * int(self.add_normal)


class NeuralLidarCBFController(nn.Module):
    def __init__(self, dynamics_model, scenarios=None, datamodule=None, experiment_suite=None, **kwargs):
        super().__init__()
        dm = self.dynamics_model = dynamics_model
        self.h_hidden_layers, self.h_hidden_size = kwargs.get("cbf_hidden_layers", 2), kwargs.get("cbf_hidden_size", 48)
        self.all_encoded_obs_dim, per = kwargs["feature_dim"], kwargs["per_feature_dim"]
        if kwargs.get("use_bn", False):
            raise NotImplementedError("use_bn=True (BatchNorm in the point network) is not built")
        S = len(dm.list_sensor)
        lk = lambda: nn.LeakyReLU(0.1)
        # construction order of the reference (pointnet.py:20-48,78-90; neural_lidar_cbf_controller.py:66-95): the same
        # torch.manual_seed gives the same parameters
        self.pc_head = nn.ModuleDict(dict(
            backbone_nn=nn.Sequential(OrderedDict([("fc11", nn.Linear(dm.point_dims + S, 64)), ("relu1", lk()), ("fc2", nn.Linear(64, 128)),
                                                   ("relu2", lk()), ("fc3", nn.Linear(128, 512)), ("relu3", lk())])),
            backbone_fc=nn.Sequential(OrderedDict([("fc2", nn.Linear(512, 256)), ("relu2", lk()), ("fc3", nn.Linear(256, per))]))))
        self.encoder = nn.ModuleDict(dict(
            encoder_nn=nn.Sequential(OrderedDict([("fc1", nn.Linear(per * S, 512)), ("relu1", lk()), ("fc2", nn.Linear(512, 256)),
                                                  ("relu2", lk()), ("fc4", nn.Linear(256, self.all_encoded_obs_dim))]))))
        h_layers = OrderedDict([("input_linear", nn.Linear(dm.n_dims + self.all_encoded_obs_dim, self.h_hidden_size)),
                                ("input_activation", lk())])
        for i in range(self.h_hidden_layers):
            h_layers[f"layer_{i}_linear"] = nn.Linear(self.h_hidden_size, self.h_hidden_size)
            h_layers[f"layer_{i}_activation"] = lk()
        h_layers["output_linear"] = nn.Linear(self.h_hidden_size, 1)
        self.h_nn = nn.Sequential(h_layers)
        self._packed = None

    @property
    def device(self):
        return torch.device(getattr(self.dynamics_model, "device", "cuda"))

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _net(self):
        if self._packed is None:
            dm = self.dynamics_model
            self._packed = ops.pack_lidar_net(self.pc_head.state_dict(), self.encoder.state_dict(), self.h_nn.state_dict(), dm.n_dims,
                                              len(dm.list_sensor), dm.point_dims, device=self.device)
        return self._packed

    def draw_sampled_index(self, n_cloud: int) -> torch.Tensor:
        """the draw of datax_to_x (arm_lidar.py:236-238): one torch.randint per sensor, shared by the whole batch"""
        dm = self.dynamics_model
        return torch.stack([torch.randint(low=0, high=n_cloud, size=(dm.ray_per_sensor, 1)).squeeze().int() for _ in dm.list_sensor])

    def _split(self, datax: torch.Tensor):
        dm = self.dynamics_model
        n, S = dm.n_dims, len(dm.list_sensor)
        bs = datax.shape[0]
        q = datax[:, :n]
        cloud = datax[:, n:-12 * S].reshape(bs, -1, dm.point_dims)
        frames = datax[:, -12 * S:].reshape(bs, S, 12)
        return q, cloud, frames

    @torch.no_grad()
    def h(self, datax: torch.Tensor, sampled_index: Optional[torch.Tensor] = None) -> torch.Tensor:
        """neural_lidar_cbf_controller.py:115-139 -> (bs, 1)"""
        q, cloud, frames = self._split(datax.to(self.device))
        idx = self.draw_sampled_index(cloud.shape[1]) if sampled_index is None else torch.as_tensor(sampled_index)
        return ops.lidar_cbf(self._net(), q, cloud, frames, idx.to(self.device)).unsqueeze(1)

    @torch.no_grad()
    def h_with_jacobian(self, datax: torch.Tensor, data_jacobian: Tuple[torch.Tensor, torch.Tensor],
                        sampled_index: Optional[torch.Tensor] = None):
        """neural_lidar_cbf_controller.py:141-219 ("pure numerical estimation") -> (h (bs,1), J (bs,1,n), t_dict)"""
        q, cloud, frames = self._split(datax.to(self.device))
        idx = self.draw_sampled_index(cloud.shape[1]) if sampled_index is None else torch.as_tensor(sampled_index)
        J_p, J_R = data_jacobian
        h, J = ops.lidar_cbf(self._net(), q, cloud, frames, idx.to(self.device), J_p.to(self.device), J_R.to(self.device),
                             delta=float(self.dynamics_model.controller_dt) / 2)
        return h.unsqueeze(1), J.unsqueeze(1), {}
