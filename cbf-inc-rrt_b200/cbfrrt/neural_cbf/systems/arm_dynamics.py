"""ArmDynamics -- mirror of neural_cbf/systems/arm_dynamics.py:17-523 with the per-state pybullet loops replaced
by batched CUDA ops.

    state  x = [theta_1 .. theta_n | o | aux]     control u = joint velocities     (f = 0, g = I)

Reference loops replaced (file:line in the reference):
    safe_mask / unsafe_mask            arm_dynamics.py:189-232  (loop :199-201, :222-224)   -> cbfrrt::masks
    complete_sample_with_observations  arm_dynamics.py:267-279  (loop :274-276)             -> cbfrrt::observe
    closed_loop_dynamics               arm_dynamics.py:474-523  (loop :497-510)             -> one torch expression
Tensors may arrive on any device; they are evaluated on ``self.device`` (the GPU) and results are returned on the
input's device, as the reference returns masks "type_as(x)" (arm_dynamics.py:197,219).
"""
import time
import warnings
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from ... import ops
from .control_affine_system import ControlAffineSystem

Scenario = Dict[str, float]
ScenarioList = List[Scenario]


class ArmDynamics(ControlAffineSystem):
    def __init__(self, nominal_params, dt: float, controller_dt: Optional[float], dis_threshold: float,
                 use_linearized_controller: bool = False, env=None, robot=None, thr_hard: float = 0.0):
        super().__init__(nominal_params, dt, controller_dt, use_linearized_controller=use_linearized_controller)
        self.dis_threshold = dis_threshold  # soft threshold of check_collision_free_soft (arm_dynamics.py:53)
        # contact threshold of check_collision_free_hard: getContactPoints is non-empty <=> distance <= thr_hard.
        # Bullet's manifold margin is not reproducible here (parity unpinned); 0 = geometric contact.
        self.thr_hard = thr_hard
        self.env = env
        self.robot = robot
        self.device = robot.device
        self.goal_state = np.array(self.robot.q0)
        self.intermediate_goals = np.array(self.robot.q0).reshape(1, -1)
        self.compute_linearized_controller(None)

    # ---- goals (arm_dynamics.py:64-74).  The reference re-solves the DARE here; K does not depend on the goal.
    def set_goal(self, q_goal):
        assert q_goal.shape == self.goal_state.shape
        assert q_goal.shape[-1] == self.intermediate_goals.shape[-1]
        self.goal_state = q_goal
        self.intermediate_goals = q_goal.reshape(1, -1)

    def set_intermediate_goals(self, q_goal):
        assert q_goal.shape[-1] == self.intermediate_goals.shape[-1]
        self.intermediate_goals = q_goal.reshape(-1, q_goal.shape[-1])

    def __str__(self):
        return f"{str(self.robot)}_Dynamics"

    @property
    def q_dims(self) -> int:
        return self.robot.body_dim

    @property
    def n_dims(self) -> int:
        return self.robot.body_dim

    @property
    def n_controls(self) -> int:
        return self.robot.body_dim

    @property
    def o_dims(self) -> int:
        return 0

    @property
    def state_aux_dims(self) -> int:
        return 0

    @property
    def state_limits(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(upper, lower) -- arm_dynamics.py:100-109"""
        return (torch.tensor(self.robot.body_range[:, 1]), torch.tensor(self.robot.body_range[:, 0]))

    @property
    def control_limits(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(upper, lower) -- arm_dynamics.py:111-119 (URDF joint velocity limits)"""
        upper_limit = torch.Tensor(self.robot.v_max)
        return (upper_limit, -upper_limit)

    def validate_params(self, params: Scenario) -> bool:
        return True

    def compute_A_matrix(self, scenario) -> np.ndarray:
        return np.zeros((self.n_dims, self.n_dims))  # arm_dynamics.py:163-167

    # ---- nominal control (arm_dynamics.py:124-161)
    def goal_tensor(self, device=None) -> torch.Tensor:
        g = self.intermediate_goals
        g = torch.from_numpy(np.asarray(g)).float() if isinstance(g, np.ndarray) else g.clone().float()
        return g.to(device if device is not None else self.device)

    def u_nominal(self, x: torch.Tensor, params: Optional[Scenario] = None) -> torch.Tensor:
        x_touse = x[:, :self.n_dims]
        K = self.K.type_as(x_touse)
        goal = self.goal_tensor(x_touse.device)
        assert goal.shape[-1] == x_touse.shape[-1]
        u = -(K @ (x_touse - goal).T).T + self.u_eq.type_as(x_touse)
        upper_u_lim, lower_u_lim = self.control_limits
        return torch.max(torch.min(u, upper_u_lim.type_as(u)), lower_u_lim.type_as(u))

    # ---- collision queries
    def _scene(self):
        return self.env.scene

    def _q_on_gpu(self, x: torch.Tensor) -> torch.Tensor:
        return x[:, :self.q_dims].to(device=self.device, dtype=torch.float32)

    def masks(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(safe, unsafe) in one launch."""
        b, s, fl, gr = self._scene()
        safe, unsafe = ops.masks(self._q_on_gpu(x), b, s, fl, gr, float(self.dis_threshold), float(self.thr_hard),
                                 self.robot.robot_id)
        return safe.bool().to(x.device), unsafe.bool().to(x.device)

    def check_collision_free_soft(self) -> bool:
        """arm_dynamics.py:169-182 at the robot's current joint state."""
        q = torch.tensor(self.robot.get_joint_position(self.robot.body_joints), dtype=torch.float32).reshape(1, -1)
        return bool(self.masks(q)[0].item())

    def check_collision_free_hard(self) -> bool:
        """arm_dynamics.py:184-187 at the robot's current joint state."""
        q = torch.tensor(self.robot.get_joint_position(self.robot.body_joints), dtype=torch.float32).reshape(1, -1)
        return not bool(self.masks(q)[1].item())

    def safe_mask(self, x):
        return self.masks(x)[0]

    def unsafe_mask(self, x):
        return self.masks(x)[1]

    def boundary_mask(self, x):
        safe, unsafe = self.masks(x)
        return torch.logical_not(torch.logical_or(safe, unsafe))

    def goal_mask(self, x, velocity_limit: bool = False):
        """arm_dynamics.py:234-262: within 0.2 rad (L2) of the goal and safe."""
        goal = torch.as_tensor(self.goal_state[:self.q_dims]).type_as(x).to(x.device)
        goal_mask = (x[:, :self.q_dims] - goal).norm(dim=1) <= 0.2
        if velocity_limit:
            goal_mask.logical_and_(x[:, 2].abs() <= 0.1)
            goal_mask.logical_and_(x[:, 3].abs() <= 0.1)
        goal_mask.logical_and_(self.safe_mask(x))
        return goal_mask

    def _get_observation_with_state(self, state):
        pass

    def complete_sample_with_observations(self, x, num_samples: int) -> torch.Tensor:
        """arm_dynamics.py:267-279; systems without an observation return x unchanged."""
        return x

    def _get_sdf(self):
        """arm_dynamics.py:281-291: shortest link-obstacle distance at the current joint state."""
        q = torch.tensor(self.robot.get_joint_position(self.robot.body_joints), dtype=torch.float32).reshape(1, -1)
        b, s, fl, gr = self._scene()
        datax = ops.observe(q.to(self.device), b, s, fl, gr, float(self.dis_threshold), float(self.thr_hard),
                            self.robot.robot_id)[2]
        return float(datax[0, self.q_dims].item())

    # ---- samplers (arm_dynamics.py:293-325)
    def sample_goal(self, num_samples: int, eps=0.02) -> torch.Tensor:
        x = torch.Tensor(num_samples, self.n_dims).uniform_(-1.0, 1.0) * eps + torch.as_tensor(self.goal_state).float()
        return self.complete_sample_with_observations(x, num_samples)

    def sample_safe(self, num_samples: int, max_tries: int = 100) -> torch.Tensor:
        x = ControlAffineSystem.sample_safe(self, num_samples, max_tries)
        return self.complete_sample_with_observations(x, num_samples)

    def sample_unsafe(self, num_samples: int, max_tries: int = 100) -> torch.Tensor:
        x = ControlAffineSystem.sample_unsafe(self, num_samples, max_tries)
        return self.complete_sample_with_observations(x, num_samples)

    def sample_boundary(self, num_samples: int, max_tries: int = 100) -> torch.Tensor:
        x = ControlAffineSystem.sample_boundary(self, num_samples, max_tries)
        return self.complete_sample_with_observations(x, num_samples)

    # ---- f = 0, g = I (arm_dynamics.py:327-363)
    def _f(self, x: torch.Tensor, params: Scenario):
        return torch.zeros((x.shape[0], self.n_dims, 1)).type_as(x)

    def _g(self, x: torch.Tensor, params: Scenario):
        if len(x.shape) > 1:
            g = torch.eye(n=self.n_dims, m=self.n_controls).unsqueeze(0).expand(x.shape[0], -1, -1)
        else:
            g = torch.eye(n=self.n_dims, m=self.n_controls)
        return g.to(x.device)

    @property
    def u_eq(self):
        return torch.zeros((1, self.n_controls))

    @property
    def goal_point(self):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return torch.tensor(self.goal_state)

    # ---- data-collection simulator (arm_dynamics.py:382-471): nominal controller + uniform control noise, held for
    # controller_period / dt steps, Euler with ``collect_dataset`` (10 sub-steps per stored state), observation
    # refreshed at every stored state.  One batched launch per stored state instead of bs pybullet loops.
    def noisy_simulator(self, x_init: torch.Tensor, num_steps: int, collect_dataset: bool = False,
                        noise_level=1e-2, use_motor_control: bool = False) -> torch.Tensor:
        assert collect_dataset
        return self.simulate(x_init, num_steps, controller=self.u_nominal, guard=self.out_of_bounds_mask,
                             noise_level=noise_level, use_motor_control=use_motor_control,
                             collect_dataset=collect_dataset)

    def simulate(self, x_init: torch.Tensor, num_steps: int, collect_dataset: bool, controller: Callable,
                 controller_period: Optional[float] = None, guard: Optional[Callable] = None,
                 noise_level: float = 0, use_motor_control: bool = False,
                 params: Optional[Scenario] = None) -> torch.Tensor:
        """-> (bs, num_steps, n + o_dims + aux) trajectories (arm_dynamics.py:400-471)."""
        assert controller.__name__ == "u_nominal"
        assert collect_dataset == True  # noqa: E712  (the reference's own guard, :431)
        init_states = self.complete_sample_with_observations(x_init, x_init.shape[0])
        x_sim = torch.zeros(x_init.shape[0], num_steps, init_states.shape[-1]).type_as(x_init)
        u = torch.zeros(x_init.shape[0], self.n_controls).type_as(x_init)
        if controller_period is None:
            controller_period = self.dt
        controller_update_freq = int(controller_period / self.dt)
        x_sim[:, 0, :] = init_states
        upper_u_lim, lower_u_lim = self.control_limits
        for tstep in range(1, num_steps):
            x_current = x_sim[:, tstep - 1, :]
            if tstep == 1 or tstep % controller_update_freq == 0:
                u = controller(x_current)
                if isinstance(u, tuple):
                    u = u[0]
                if noise_level > 0:
                    u += torch.mul(torch.Tensor(x_init.shape[0], self.n_controls).uniform_(-noise_level, noise_level)
                                   .type_as(x_init), upper_u_lim.type_as(x_init))
                    u = torch.max(torch.min(u, upper_u_lim.type_as(u)), lower_u_lim.type_as(u))
            x_sim[:, tstep, :] = self.closed_loop_dynamics(x_current, u, use_motor_control=use_motor_control,
                                                           collect_dataset=collect_dataset)
        return x_sim

    # ---- Euler step (arm_dynamics.py:474-523).  NOTE: returns x_next, not xdot (:480-485).
    def closed_loop_dynamics(self, x: torch.Tensor, u: torch.Tensor, collect_dataset: bool = False,
                             use_motor_control: bool = False, params: Optional[Scenario] = None,
                             update_observation: bool = True, return_time: bool = False):
        if use_motor_control:
            raise NotImplementedError("motor control needs a physics step (arm_dynamics.py:499-506); out of scope")
        t0 = time.time()
        batch_size = x.shape[0]
        width = self.n_dims + self.o_dims + self.state_aux_dims if self.o_dims else self.n_dims
        x_next = torch.zeros((batch_size, width)).type_as(x)
        step = 10 if collect_dataset else 1
        x_next[:, :self.n_dims] = u.type_as(x) * self.dt * step + x[:, :self.n_dims]
        t1 = time.time()
        if update_observation:
            x_next = self.complete_sample_with_observations(x_next[:, :self.n_dims], batch_size)
        t2 = time.time()
        if return_time:
            return x_next, [t1 - t0, t2 - t1]
        return x_next
