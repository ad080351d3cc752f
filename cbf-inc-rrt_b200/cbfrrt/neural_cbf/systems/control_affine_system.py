"""ControlAffineSystem -- mirror of neural_cbf/systems/control_affine_system.py:23-584 (the members the hot path
and its callers use; plotting / robust multi-scenario Lyapunov synthesis are out of scope)."""
from abc import ABC, abstractmethod
from typing import Callable, Optional, Tuple

import numpy as np
import torch

from .utils import Scenario, ScenarioList, lqr, continuous_lyap


class ControlAffineSystem(ABC):
    """dx/dt = f(x) + g(x) u"""

    def __init__(self, nominal_params: Scenario, dt: float = 0.01, controller_dt: Optional[float] = None,
                 use_linearized_controller: bool = True, scenarios: Optional[ScenarioList] = None):
        super().__init__()
        if not self.validate_params(nominal_params):
            raise ValueError(f"Parameters not valid: {nominal_params}")
        self.nominal_params = nominal_params
        assert dt > 0.0
        self.dt = dt
        self.controller_dt = controller_dt
        if use_linearized_controller:
            self.compute_linearized_controller(scenarios)

    # ---- linearisation (control_affine_system.py:79-157)
    def compute_A_matrix(self, scenario):
        raise NotImplementedError

    def compute_B_matrix(self, scenario) -> np.ndarray:
        if scenario is None:
            scenario = self.nominal_params
        B = self._g(self.goal_point, scenario).squeeze().cpu().numpy()
        return np.reshape(B, (self.n_dims, self.n_controls))

    def linearized_ct_dynamics_matrices(self, scenario=None):
        return self.compute_A_matrix(scenario), self.compute_B_matrix(scenario)

    def linearized_dt_dynamics_matrices(self, scenario=None):
        Act, Bct = self.linearized_ct_dynamics_matrices(scenario)
        return np.eye(self.n_dims) + self.controller_dt * Act, self.controller_dt * Bct

    def compute_linearized_controller(self, scenarios: Optional[ScenarioList] = None):
        if scenarios is None:
            scenarios = [self.nominal_params]
        Acl_list = []
        for s in scenarios:
            Act, Bct = self.linearized_ct_dynamics_matrices(s)
            A, B = self.linearized_dt_dynamics_matrices(s)
            Q, R = np.eye(self.n_dims), np.eye(self.n_controls)
            K_np = lqr(A, B, Q, R)
            self.K = torch.tensor(K_np)
            Acl_list.append(Act - Bct @ K_np)
        if len(scenarios) > 1:
            raise NotImplementedError("robust multi-scenario Lyapunov synthesis needs cvxpy (utils.py:74); out of scope")
        self.P = torch.tensor(continuous_lyap(Acl_list[0], Q))

    # ---- abstract interface
    @abstractmethod
    def validate_params(self, params: Scenario) -> bool: ...

    @property
    @abstractmethod
    def n_dims(self) -> int: ...

    @property
    @abstractmethod
    def state_limits(self) -> Tuple[torch.Tensor, torch.Tensor]: ...

    @property
    @abstractmethod
    def control_limits(self) -> Tuple[torch.Tensor, torch.Tensor]: ...

    @property
    def intervention_limits(self):
        upper_limit, lower_limit = self.control_limits
        return (upper_limit, lower_limit)

    def out_of_bounds_mask(self, x: torch.Tensor) -> torch.Tensor:
        """control_affine_system.py:200-216"""
        upper_lim, lower_lim = self.state_limits
        upper_lim, lower_lim = upper_lim.to(x.device), lower_lim.to(x.device)
        m = torch.zeros_like(x[:, 0], dtype=torch.bool)
        for i_dim in range(x.shape[-1]):
            m.logical_or_(x[:, i_dim] >= upper_lim[i_dim])
            m.logical_or_(x[:, i_dim] <= lower_lim[i_dim])
        return m

    @abstractmethod
    def safe_mask(self, x: torch.Tensor) -> torch.Tensor: ...

    @abstractmethod
    def unsafe_mask(self, x: torch.Tensor) -> torch.Tensor: ...

    def failure(self, x: torch.Tensor) -> torch.Tensor:
        return self.unsafe_mask(x)

    def boundary_mask(self, x: torch.Tensor) -> torch.Tensor:
        """control_affine_system.py:254-268"""
        return torch.logical_not(torch.logical_or(self.safe_mask(x), self.unsafe_mask(x)))

    def goal_mask(self, x: torch.Tensor) -> torch.Tensor:
        return (x - self.goal_point.to(x.device)).norm(dim=-1) <= 0.1

    @property
    def goal_point(self):
        return torch.zeros((1, self.n_dims))

    @property
    def u_eq(self):
        return torch.zeros((1, self.n_controls))

    def K_on(self, device) -> torch.Tensor:
        """float32 copy of the LQR gain on ``device``, uploaded once per gain (the planners call the filter thousands of
        times per problem with the same K)."""
        key = (id(self.K), self.K._version, str(device))
        if getattr(self, "_K_dev_key", None) != key:
            self._K_dev = self.K.to(device, torch.float32).contiguous()
            self._K_dev_key = key
        return self._K_dev

    # ---- samplers (control_affine_system.py:291-354)
    def sample_state_space(self, num_samples: int) -> torch.Tensor:
        x_max, x_min = self.state_limits
        x = torch.Tensor(num_samples, self.n_dims).uniform_(0.0, 1.0)
        # column by column in the reference (control_affine_system.py:296-298); the broadcast form does the same float32
        # operations per element
        return x * (x_max - x_min).reshape(1, -1).type_as(x) + x_min.reshape(1, -1).type_as(x)

    def sample_with_mask(self, num_samples: int, mask_fn: Callable[[torch.Tensor], torch.Tensor], max_tries: int = 100):
        samples = self.sample_state_space(max(2 * num_samples, 100))
        mask_result = mask_fn(samples)
        for _ in range(max_tries):
            if mask_result.sum().item() >= num_samples:
                return samples[mask_result][:num_samples]
            violations = torch.logical_not(mask_result)
            samples[violations] = self.sample_state_space(int(violations.sum().item()))
            mask_result = mask_fn(samples)
        raise RuntimeWarning(f"unable to collect enough points within {max_tries} tries.")

    def sample_safe(self, num_samples: int, max_tries: int = 100) -> torch.Tensor:
        return self.sample_with_mask(num_samples, self.safe_mask, max_tries)

    def sample_unsafe(self, num_samples: int, max_tries: int = 100) -> torch.Tensor:
        return self.sample_with_mask(num_samples, self.unsafe_mask, max_tries)

    def sample_goal(self, num_samples: int, max_tries: int = 100) -> torch.Tensor:
        return self.sample_with_mask(num_samples, self.goal_mask, max_tries)

    def sample_boundary(self, num_samples: int, max_tries: int = 100) -> torch.Tensor:
        return self.sample_with_mask(num_samples, self.boundary_mask, max_tries)

    # ---- dynamics (control_affine_system.py:356-402)
    def control_affine_dynamics(self, x: torch.Tensor, params: Optional[Scenario] = None):
        assert x.ndim == 2
        if params is None:
            params = self.nominal_params
        return self._f(x, params), self._g(x, params)

    def closed_loop_dynamics(self, x: torch.Tensor, u: torch.Tensor, params: Optional[Scenario] = None) -> torch.Tensor:
        f, g = self.control_affine_dynamics(x, params=params)
        xdot = f + torch.bmm(g, u.unsqueeze(-1))
        return xdot.view(x.shape)

    @abstractmethod
    def _f(self, x: torch.Tensor, params: Scenario) -> torch.Tensor: ...

    @abstractmethod
    def _g(self, x: torch.Tensor, params: Scenario) -> torch.Tensor: ...

    def u_nominal(self, x: torch.Tensor, params: Optional[Scenario] = None) -> torch.Tensor:
        """control_affine_system.py:547-574"""
        K = self.K.type_as(x)
        goal = self.goal_point.squeeze().type_as(x)
        u = -(K @ (x - goal).T).T + self.u_eq.type_as(x)
        upper_u_lim, lower_u_lim = self.control_limits
        return torch.max(torch.min(u, upper_u_lim.type_as(x)), lower_u_lim.type_as(x))
