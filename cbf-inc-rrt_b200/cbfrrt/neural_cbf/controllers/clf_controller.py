"""CLFController -- mirror of neural_cbf/controllers/clf_controller.py:29-535 (inference path).

The reference builds a CvxpyLayer (:82-139) and solves one SCS conic program per row (:354-404).  Here the same
QP     min ||u - u_ref||^2 + p * sum r_i   s.t.  Lf_i + Lg_i u + lambda V - r_i <= 0,  r_i >= 0,  lo <= u <= hi
is solved exactly for the whole batch: cbfrrt::qp (KKT closed form) for n_scenarios == 1 -- the hot path of every
shipped entry point -- and cbfrrt::qp_multi (dual active-set coordinate solver, up to 4 scenarios) otherwise.
"""
import time
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from ... import ops
from .controller import Controller


class CLFController(Controller):
    def __init__(self, dynamics_model, scenarios, experiment_suite=None, clf_lambda: float = 1.0,
                 clf_relaxation_penalty: float = 50.0, controller_period: float = 0.01, **kwargs):
        super(CLFController, self).__init__(dynamics_model=dynamics_model, experiment_suite=experiment_suite,
                                            controller_period=controller_period)
        self.scenarios = scenarios
        self.n_scenarios = len(scenarios)
        self.experiment_suite = experiment_suite
        self.clf_lambda = clf_lambda
        if "safe_level" in kwargs:
            self.safe_level = kwargs["safe_level"]
            self.unsafe_level = kwargs["unsafe_level"]
        self.clf_relaxation_penalty = clf_relaxation_penalty
        # multi-scenario solver budget (cbfrrt_qp_multi): iterations and control-space tolerance
        self.qp_max_sweeps = int(kwargs.get("qp_max_sweeps", 100))
        self.qp_tol = float(kwargs.get("qp_tol", 1e-9))

    @property
    def device(self):
        return self.dynamics_model.device

    def V(self, x: torch.Tensor) -> torch.Tensor:
        V, _ = self.V_with_jacobian(x)
        return V

    def V_with_jacobian(self, x: torch.Tensor, data_jacobian: tuple = ()):
        """clf_controller.py:146-167: V = 0.5 x^T P x, JV = P x.  (3-tuple like the neural subclasses so that
        V_with_lie_derivatives can call either.)"""
        n = self.dynamics_model.n_dims
        P = self.dynamics_model.P.type_as(x).reshape(1, n, n)
        V = 0.5 * F.bilinear(x, x, P).squeeze().reshape(x.shape[0])
        JV = F.linear(x, P.reshape(n, n)).reshape(x.shape[0], 1, n)
        return V, JV, {}

    def V_with_lie_derivatives(self, x: torch.Tensor, data_jacobian: tuple = (), scenarios=None):
        """clf_controller.py:169-218 -> (V, Lf_V (bs,S,1), Lg_V (bs,S,n), t_dict)"""
        t_dict = {}
        t0 = time.time()
        if scenarios is None:
            scenarios = self.scenarios
        n_scenarios = len(scenarios)
        V, gradV, jacob_tdict = self.V_with_jacobian(x, data_jacobian=data_jacobian)
        t_dict.update(jacob_tdict)
        t_dict["V_w_Jacobian"] = time.time() - t0
        t1 = time.time()
        bs = x.shape[0]
        Lf_V = torch.zeros(bs, n_scenarios, 1).type_as(gradV)
        Lg_V = torch.zeros(bs, n_scenarios, self.dynamics_model.n_controls).type_as(gradV)
        for i in range(n_scenarios):
            f, g = self.dynamics_model.control_affine_dynamics(x[:, :self.dynamics_model.n_dims], params=scenarios[i])
            Lf_V[:, i, :] = torch.bmm(gradV, f.type_as(gradV).to(gradV.device)).squeeze(1)
            Lg_V[:, i, :] = torch.bmm(gradV, g.type_as(gradV).to(gradV.device)).squeeze(1)
        t_dict["lie_derivative"] = time.time() - t1
        return V, Lf_V, Lg_V, t_dict

    def u_reference(self, x: torch.Tensor) -> torch.Tensor:
        return self.dynamics_model.u_nominal(x)

    def _solve_CLF_QP_exact(self, x, u_ref, V, Lf_V, Lg_V, relaxation_penalty: float, requires_grad: bool = False):
        """Batched exact solve on the GPU (replaces _solve_CLF_QP_cvxpylayers, clf_controller.py:354-404).  With
        ``requires_grad`` the solve is differentiable w.r.t. V, Lf_V, Lg_V and u_ref like the reference's CvxpyLayer
        (closed-form active-set sensitivities: cbfrrt::qp_backward, cbfrrt::qp_multi_backward)."""
        dev = self.device
        upper, lower = self.dynamics_model.control_limits
        if self.n_scenarios != 1:
            a = Lg_V.to(dev, torch.float32)
            c = (Lf_V[:, :, 0] + self.clf_lambda * V.reshape(-1, 1)).to(dev, torch.float32)
            solve = ops.qp_multi_diff if requires_grad else ops.qp_multi
            u, r = solve(a, c, u_ref.to(dev, torch.float32), lower.to(dev, torch.float32), upper.to(dev, torch.float32),
                         float(relaxation_penalty), self.qp_max_sweeps, self.qp_tol)[:2]
            return u.to(x.device).type_as(x), r.to(x.device).type_as(x)
        a = Lg_V[:, 0, :].to(dev, torch.float32)
        c = (Lf_V[:, 0, 0] + self.clf_lambda * V.reshape(-1)).to(dev, torch.float32)
        solve = ops.qp_diff if requires_grad else ops.qp
        u, r = solve(a, c, u_ref.to(dev, torch.float32), lower.to(dev, torch.float32), upper.to(dev, torch.float32),
                     float(relaxation_penalty))
        return u.to(x.device).type_as(x), r.reshape(-1, 1).to(x.device).type_as(x)

    # ---- the reference's other back-ends, same signatures and return conventions, all on the exact batched solver
    def _solve_CLF_QP_cvxpylayers(self, x, u_ref, V, Lf_V, Lg_V, relaxation_penalty: float):
        """clf_controller.py:354-404 (differentiable layer; penalty clamped to 1e6, :380)"""
        return self._solve_CLF_QP_exact(x, u_ref, V, Lf_V, Lg_V, relaxation_penalty,
                                        requires_grad=torch.is_grad_enabled() and any(
                                            t.requires_grad for t in (u_ref, V, Lf_V, Lg_V)))

    @torch.no_grad()
    def _solve_CLF_QP_gurobi(self, x, u_ref, V, Lf_V, Lg_V, relaxation_penalty: float):
        """clf_controller.py:226-352: the same relaxed QP per row.  Rows with a NaN / inf in x, Lf_V or Lg_V are skipped
        (u = 0, r = 0) like the reference's (:281-290).  Difference: the penalty is clamped at 1e6 by the library, so a
        hard constraint (penalty = inf) that is infeasible inside the control box comes back relaxed (r > 0) where
        Gurobi reports infeasibility (r = NaN, :325-331)."""
        bad = ~(torch.isfinite(x).all(dim=1) & torch.isfinite(Lg_V.reshape(x.shape[0], -1)).all(dim=1)
                & torch.isfinite(Lf_V.reshape(x.shape[0], -1)).all(dim=1))
        clean = lambda t: torch.where(bad.reshape(-1, *[1] * (t.dim() - 1)), torch.zeros_like(t), t)
        u, r = self._solve_CLF_QP_exact(x, clean(u_ref), clean(V), clean(Lf_V), clean(Lg_V), min(relaxation_penalty, 1e6))
        u[bad] = 0.0
        r[bad] = 0.0
        return u, r

    @torch.no_grad()
    def _solve_CLF_QP_qpsolver(self, x, u_ref, V, Lf_V, Lg_V, relaxation_penalty: float, use_constraint: bool = False):
        """clf_controller.py:405-459 (OSQP through qpsolvers, one scenario).  ``use_constraint=False``: the penalty
        form min ||u - u_ref||^2 + p Lg.u over the box, i.e. u = clip(u_ref - p Lg / 2) (:431-435), r = 0.
        ``use_constraint=True``: the hard-constrained QP (:423-429); when any row of the batch is infeasible inside the
        box the reference's solver fails on the stacked problem and the whole batch falls back to the penalty form
        (:449-452) -- reproduced here from the relaxation the exact solver reports at p = 1e6."""
        assert self.n_scenarios == 1
        upper, lower = self.dynamics_model.control_limits
        if use_constraint:
            u, r = self._solve_CLF_QP_exact(x, u_ref, V, Lf_V, Lg_V, 1e6)
            scale = 1.0 + Lf_V.reshape(x.shape[0], -1).abs().sum(1) + Lg_V.reshape(x.shape[0], -1).abs().sum(1)
            if bool((r.reshape(-1) <= 1e-6 * scale.to(r.device)).all()):
                return u, torch.zeros_like(V)
        u = u_ref - 0.5 * relaxation_penalty * Lg_V[:, 0, :].type_as(u_ref)
        u = torch.max(torch.min(u, upper.type_as(u).to(u.device)), lower.type_as(u).to(u.device))
        return u.type_as(x), torch.zeros_like(V)

    def solve_CLF_QP(self, x, relaxation_penalty: Optional[float] = None, u_ref: Optional[torch.Tensor] = None,
                     requires_grad: bool = False, strict_solution: bool = False) -> Tuple[tuple, dict]:
        """clf_controller.py:461-528 -> ((u, r), t_dict)"""
        t_dict = {}
        a = time.time()
        V, Lf_V, Lg_V, lie_t_dict = self.V_with_lie_derivatives(x)
        t_dict.update(lie_t_dict)
        t_dict["nn_forward"] = time.time() - a
        b = time.time()
        if u_ref is not None:
            assert u_ref.shape[0] == x.shape[0], f"u_ref must have {x.shape[0]} rows, but got {u_ref.shape[0]}"
            assert u_ref.shape[1] == self.dynamics_model.n_controls
        else:
            u_ref = self.u_reference(x)
        if relaxation_penalty is None:
            relaxation_penalty = self.clf_relaxation_penalty
        sol = self._solve_CLF_QP_exact(x, u_ref, V, Lf_V, Lg_V, relaxation_penalty, requires_grad=requires_grad)
        t_dict["QP"] = time.time() - b
        return sol, t_dict

    def u(self, x):
        """clf_controller.py:530-535 -> (u, t_dict)"""
        u, t_dict = self.solve_CLF_QP(x)
        return u[0], t_dict
