"""ArmMindis -- mirror of neural_cbf/systems/arm_mindis.py:16-90.

Observation o = distance to the closest obstacle (1 value), aux = do/dq (n values):
datax = [q | o | do/dq], computed for the whole batch by cbfrrt::observe instead of the per-state
getClosestPoints / calculateJacobian calls (arm_mindis.py:56-90, arm_dynamics.py:274-276).
"""
from typing import Optional

import numpy as np
import torch

from ... import ops
from .arm_dynamics import ArmDynamics, Scenario


class ArmMindis(ArmDynamics):
    def __init__(self, nominal_params: Scenario, dt: float, controller_dt: Optional[float], dis_threshold: float,
                 env=None, robot=None, thr_hard: float = 0.0):
        super().__init__(nominal_params, dt, controller_dt, dis_threshold=dis_threshold,
                         use_linearized_controller=False, env=env, robot=robot, thr_hard=thr_hard)

    def __str__(self):
        return f"{str(self.robot)}_mindis_Dynamics"

    @property
    def o_dims(self) -> int:
        return 1

    @property
    def state_aux_dims(self) -> int:
        return self.robot.body_dim

    def observe(self, x: torch.Tensor):
        """(safe, unsafe, datax) for the batch in one launch."""
        b, s, fl, gr = self._scene()
        safe, unsafe, datax = ops.observe(self._q_on_gpu(x), b, s, fl, gr, float(self.dis_threshold),
                                          float(self.thr_hard), self.robot.robot_id)
        return safe.bool().to(x.device), unsafe.bool().to(x.device), datax.to(x.device).type_as(x)

    def complete_sample_with_observations(self, x, num_samples: int) -> torch.Tensor:
        """arm_dynamics.py:267-279 -> (num_samples, 2n+1)"""
        return self.observe(x[:num_samples])[2]

    def _get_observation_with_state(self, state):
        """arm_mindis.py:71-90: np.concatenate(([o], do_dq)) for one state."""
        x = torch.as_tensor(np.asarray(state), dtype=torch.float32).reshape(1, -1)
        self.robot.set_joint_position(self.robot.body_joints, x[0, :self.q_dims])
        datax = self.observe(x)[2]
        return datax[0, self.q_dims:].double().cpu().numpy()
