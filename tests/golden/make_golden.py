#!/usr/bin/env python3
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN PYTHON (dev container only).

The reference (/root/reference) cannot be imported as shipped: pybullet, pinocchio, cvxpy/cvxpylayers,
pytorch_lightning, matplotlib ... are not installed and `neural_cbf/controllers/__init__.py` imports a missing
`baselines` package.  Everything on the hot path that is *pure torch / numpy / scipy* can still
be run unmodified if the absent third-party modules are replaced by inert stubs.  This script does exactly
that and records inputs/outputs of the reference's own functions:

  neural_cbf/systems/control_affine_system.py:125-157 + utils.py:18-49   LQR gain K (scipy DARE)
  neural_cbf/systems/arm_dynamics.py:124-161                             u_nominal (+ set_goal / set_intermediate_goals)
  neural_cbf/controllers/neural_mindis_cbf_controller.py:41-53,71-102    h, h_with_jacobian (autograd)
  neural_cbf/controllers/clf_controller.py:169-218                       V_with_lie_derivatives
  neural_cbf/controllers/clf_controller.py:354-404,461-535               solve_CLF_QP / u  (parameter assembly)
  neural_cbf/systems/arm_dynamics.py:474-523                             closed_loop_dynamics (Euler step, no refresh)
  neural_cbf/systems/control_affine_system.py:200-216                    out_of_bounds_mask

What is stubbed and therefore NOT pinned by these vectors:
  * pybullet: the `robot`/`env` objects handed to ArmMindis are fakes exposing only joint limits, velocity
    limits and q0 (numbers copied from the URDF via spec/robot_models.json).  No FK / distance comes from here.
  * cvxpylayers/SCS: `CvxpyLayer.__call__` is replaced by an independent float64 solve of the same QP with
    scipy SLSQP (tight tolerances).  It pins the QP *definition and parameter plumbing*, not SCS's iterates.
  * pytorch_lightning.LightningModule -> torch.nn.Module.

Output: tests/golden/ref_controller_{panda,magician}.npz  (committed; the GPU box has no /root/reference).
Run:   python tests/golden/make_golden.py
"""
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CBFRRT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


# ----------------------------------------------------------------------------- inert stubs
class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _Dummy(metaclass=_DummyMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:  # used as a decorator
            return a[0]
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()

    def __getitem__(self, k):
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return True

    def _op(self, *a):
        return _Dummy()

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __matmul__ = __rmatmul__ = _op
    __le__ = __ge__ = __lt__ = __gt__ = __neg__ = __lshift__ = __rshift__ = _op


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _LightningModuleStandIn(torch.nn.Module):
    """pl.LightningModule stand-in: an nn.Module whose __init__ forwards the keyword arguments along the MRO to
    the controller bases, as Lightning 1.3's cooperative __init__ does."""

    def __init__(self, *args, **kwargs):
        torch.nn.Module.__init__(self)
        super(torch.nn.Module, self).__init__(*args, **kwargs)


STUB_ROOTS = ["pybullet", "pybullet_utils", "pybullet_data", "pinocchio", "cvxpy", "cvxpylayers", "gurobipy",
              "qpsolvers", "pytorch_lightning", "matplotlib", "seaborn", "neural_cbf.controllers.baselines"]


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        for r in STUB_ROOTS:
            if fullname == r or fullname.startswith(r + "."):
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        if module.__name__ == "pytorch_lightning":
            module.LightningModule = _LightningModuleStandIn
        if module.__name__ == "cvxpylayers.torch":
            module.CvxpyLayer = SlsqpLayer


# ----------------------------------------------------------------------------- stand-in for cvxpylayers+SCS
QP_LIMITS = {}


class SlsqpLayer:
    """Solves  min ||u-u_ref||^2 + p r  s.t. Lf + Lg.u + lam*V - r <= 0, r >= 0, lo <= u <= hi   per row,
    with the parameter ORDER the reference passes (clf_controller.py:383-391): Lf, Lg, V, u_ref, penalty."""

    def __init__(self, problem, variables=None, parameters=None):
        pass

    def __call__(self, *params, solver_args=None):
        from scipy.optimize import minimize
        Lf, Lg, V, u_ref, pen = [p.detach().double().numpy() for p in params]
        lam, lo, hi = QP_LIMITS["lam"], QP_LIMITS["lo"], QP_LIMITS["hi"]
        bs, n = u_ref.shape
        U = np.zeros((bs, n))
        R = np.zeros((bs, 1))
        p = float(pen[0])
        for b in range(bs):
            a, c, ur = Lg[b], float(Lf[b, 0] + lam * V[b, 0]), u_ref[b]
            # scale r so that the problem is well conditioned for SLSQP (r = s / p)
            f = lambda z: np.sum((z[:n] - ur) ** 2) + z[n]
            g = lambda z: np.concatenate([2 * (z[:n] - ur), [1.0]])
            cons = [{"type": "ineq", "fun": lambda z: -(c + a @ z[:n] - z[n] / p),
                     "jac": lambda z: np.concatenate([-a, [1.0 / p]])}]
            bnds = [(lo[i], hi[i]) for i in range(n)] + [(0.0, None)]
            best = None
            for z0 in (np.concatenate([ur, [0.0]]), np.concatenate([np.clip(ur - a, lo, hi), [max(0.0, c) * p]])):
                res = minimize(f, z0, jac=g, bounds=bnds, constraints=cons, method="SLSQP",
                               options=dict(ftol=1e-15, maxiter=500))
                if best is None or res.fun < best.fun:
                    best = res
            U[b] = best.x[:n]
            R[b, 0] = best.x[n] / p
        return torch.tensor(U), torch.tensor(R)


# ----------------------------------------------------------------------------- fake pybullet-backed objects
class _FakeBullet:
    def __init__(self, model):
        self.m = model

    def getJointInfo(self, robot_id, j):
        info = [None] * 13
        l = self.m["links"][j]
        info[8], info[9], info[11] = l["lower"], l["upper"], l["velocity"]
        return tuple(info)


class _FakeRobot:
    def __init__(self, model):
        self.robotId = 0
        self.body_joints = model["body_joints"]
        self.body_dim = model["n"]
        self.ee_joints = model["ee_joints"]
        self.ee_dim = len(model["ee_joints"])
        self.n_joints = model["n_links"]
        self.body_range = np.array([(model["q_lower"][i], model["q_upper"][i]) for i in range(model["n"])])
        self.q0 = np.array(model["q0"])

    def set_joint_position(self, joints, q):       # pybullet resetJointState: irrelevant to the arithmetic recorded here
        pass


class _FakeEnv:
    def __init__(self, model):
        self.p = _FakeBullet(model)
        self.obstacle_ids = []


def main():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REF)
    from neural_cbf.systems import ArmMindis  # noqa: E402  (reference code)
    from neural_cbf.controllers.neural_mindis_cbf_controller import NeuralMindisCBFController  # noqa: E402

    models = json.load(open(os.path.join(ROOT, "spec", "robot_models.json")))
    for name in ("panda", "magician"):
        m = models[name]
        n = m["n"]
        env, robot = _FakeEnv(m), _FakeRobot(m)
        dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=robot)
        K = dyn.K.numpy().copy()  # reference LQR (scipy DARE)
        up, lo_ = dyn.control_limits
        lam, pen = 0.1, 5000.0
        QP_LIMITS.update(lam=lam, lo=lo_.double().numpy(), hi=up.double().numpy())
        torch.manual_seed(1)
        ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, cbf_alpha=lam, cbf_relaxation_penalty=pen,
                                         safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48,
                                         use_neural_actor=False)
        sd = {k: v.detach().numpy().copy() for k, v in ctrl.h_nn.state_dict().items()}

        g = torch.Generator().manual_seed(7)
        B = 96
        qlo, qhi = torch.tensor(m["q_lower"]).float(), torch.tensor(m["q_upper"]).float()
        q = qlo + (qhi - qlo) * torch.rand(B, n, generator=g)
        o = -0.1 + 0.4 * torch.rand(B, 1, generator=g)
        dodq = 0.5 * torch.randn(B, n, generator=g)
        datax = torch.cat([q, o, dodq], dim=1)
        # make a third of the rows "QP active": scale the observation gradient up so that Lg.u_ref + lam*h > 0
        datax[::3, n + 1:] *= 6.0

        out = {}
        # (1) shared goal (set_goal) and (2) per-row goals (set_intermediate_goals, batch_tsa.py:199)
        goal_shared = np.array(m["q0"], dtype=np.float64)
        goal_rows = (qlo + (qhi - qlo) * torch.rand(B, n, generator=g)).numpy().astype(np.float64)
        for tag, setter, goal in (("shared", dyn.set_goal, goal_shared), ("rows", dyn.set_intermediate_goals, goal_rows)):
            setter(goal)
            h = ctrl.h(datax)
            h2, J, _ = ctrl.h_with_jacobian(datax)
            V, Lf, Lg, _ = ctrl.V_with_lie_derivatives(datax)
            u_ref = dyn.u_nominal(datax)
            (u, r), _ = ctrl.solve_CLF_QP(datax)
            u2, _ = ctrl.u(datax)
            assert torch.equal(h, h2) and torch.equal(u, u2)
            # small penalty -> relaxation branch (mu = p, r > 0); huge penalty -> the min(p, 1e6) clamp (:380)
            (u_s, r_s), _ = ctrl.solve_CLF_QP(datax, relaxation_penalty=0.05)
            (u_b, r_b), _ = ctrl.solve_CLF_QP(datax, relaxation_penalty=1e9)
            out.update({f"{tag}_u_pen0.05": u_s.numpy(), f"{tag}_r_pen0.05": r_s.numpy(),
                        f"{tag}_u_pen1e9": u_b.numpy(), f"{tag}_r_pen1e9": r_b.numpy()})
            out.update({f"{tag}_goal": goal, f"{tag}_h": h.detach().numpy(), f"{tag}_J": J.detach().numpy(),
                        f"{tag}_V": V.detach().numpy(), f"{tag}_Lf": Lf.detach().numpy(),
                        f"{tag}_Lg": Lg.detach().numpy(), f"{tag}_u_ref": u_ref.detach().numpy(),
                        f"{tag}_u": u.detach().numpy(), f"{tag}_r": r.detach().numpy()})
        # Euler step of closed_loop_dynamics without the observation refresh (arm_dynamics.py:474-523, pure torch once
        # set_joint_position is a no-op) and the out-of-bounds mask (control_affine_system.py:200-216)
        dyn.set_goal(goal_shared)
        x_next = dyn.closed_loop_dynamics(datax, torch.tensor(out["shared_u"]).float(), update_observation=False)
        x_wide = datax.clone()
        x_wide[::5, 0] += 10.0
        x_wide = x_wide[:, :n]        # the reference indexes the limits with the column index: q columns only
        out.update(cld_x_next=x_next.detach().numpy(), oob_x=x_wide.numpy(), oob_mask=dyn.out_of_bounds_mask(x_wide).numpy())
        # state limits come back as (upper, lower): arm_dynamics.py:107-109
        s_up, s_lo = dyn.state_limits
        np.savez(os.path.join(HERE, f"ref_controller_{name}.npz"), datax=datax.numpy(), K=K,
                 u_upper=up.numpy(), u_lower=lo_.numpy(), x_upper=s_up.numpy(), x_lower=s_lo.numpy(),
                 lam=lam, penalty=pen, dt=dyn.dt, controller_dt=dyn.controller_dt,
                 **{"sd_" + k: v for k, v in sd.items()}, **out)
        active = int((out["shared_r"] > 1e-9).sum()), int((np.abs(out["shared_u"] - out["shared_u_ref"]).max(1) > 1e-6).sum())
        print(name, "K00", K[0, 0], "params", sum(v.size for v in sd.values()), "relaxed rows / filtered rows", active,
              "relaxed rows at p=0.05:", int((out["shared_r_pen0.05"] > 1e-9).sum()))


if __name__ == "__main__":
    main()
