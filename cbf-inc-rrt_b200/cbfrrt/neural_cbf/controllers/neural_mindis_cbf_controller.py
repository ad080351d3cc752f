"""NeuralMindisCBFController -- mirror of neural_cbf/controllers/neural_mindis_cbf_controller.py:21-102.

The barrier network h_nn is built exactly like the reference (:41-53) so checkpoints' ``h_nn.*`` tensors load
unchanged, but it is never executed by torch: its weights are packed once into the library's layout and every
inference method is ONE launch of the fused CUDA op cbfrrt::cbf_filter, which evaluates

    h (:71-81)  ->  Jh by exact reverse mode (:93-96)  ->  J = Jh_q + Jh_o do/dq (:99-101)  ->  Lf = 0, Lg = J
    (clf_controller.py:169-218, f = 0, g = I)  ->  u_ref = clamp(-K (q - goal)) (arm_dynamics.py:124-161)  ->
    exact CBF-QP (clf_controller.py:82-139,354-404)

for the whole batch.  ``u`` keeps the reference's ``(u, t_dict)`` return convention (tsa.py:232-240 depends on it).
"""
import time
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from .neural_obs_cbf_controller import NeuralObsCBFController


class NeuralMindisCBFController(NeuralObsCBFController):
    def __init__(self, dynamics_model, scenarios, datamodule=None, experiment_suite=None, **kwargs):
        super(NeuralMindisCBFController, self).__init__(dynamics_model=dynamics_model, scenarios=scenarios,
                                                        datamodule=datamodule, experiment_suite=experiment_suite,
                                                        **kwargs)
        self.n_dims_extended = self.dynamics_model.n_dims + self.dynamics_model.o_dims
        num_h_inputs = self.n_dims_extended
        self.h_layers: "OrderedDict[str, nn.Module]" = OrderedDict()
        self.h_layers["input_linear"] = nn.Linear(num_h_inputs, self.h_hidden_size)
        self.h_layers["input_activation"] = nn.ReLU()
        for i in range(self.h_hidden_layers):
            self.h_layers[f"layer_{i}_linear"] = nn.Linear(self.h_hidden_size, self.h_hidden_size)
            self.h_layers[f"layer_{i}_activation"] = nn.ReLU()
        self.h_layers["output_linear"] = nn.Linear(self.h_hidden_size, 1)
        self.h_nn = nn.Sequential(self.h_layers)
        self._packed = None
        self._packed_key = None
        self._consts = None

    # ------------------------------------------------------------------ device-side constants
    def _weights_key(self):
        ps = getattr(self, "_param_list", None)
        if ps is None:
            ps = self._param_list = list(self.h_nn.parameters())
        return tuple((p.data_ptr(), p._version) for p in ps)

    def _weights(self) -> torch.Tensor:
        key = self._weights_key()
        if self._packed is None or key != self._packed_key:
            self._packed = ops.pack_weights(self.h_nn.state_dict(), self.dynamics_model.n_dims, self.h_hidden_size,
                                            self.h_hidden_layers, device=self.device)
            self._packed_key = key
        return self._packed

    def _weights_host(self):
        """the packed network as a numpy buffer, for the host-pointer entry points (cbfrrt_*_host)"""
        key = self._weights_key()
        if getattr(self, "_packed_np", None) is None or key != getattr(self, "_packed_np_key", None):
            from ...workloads import pack_weights_np
            self._packed_np = pack_weights_np(self.h_nn.state_dict(), self.dynamics_model.n_dims, self.h_hidden_size,
                                              self.h_hidden_layers)
            self._packed_np_key = key
        return self._packed_np

    def _limits_host(self):
        if getattr(self, "_consts_np", None) is None:
            upper, lower = self.dynamics_model.control_limits
            self._consts_np = (lower.detach().cpu().numpy().astype("float32"), upper.detach().cpu().numpy().astype("float32"))
        return self._consts_np

    def _limits(self):
        if self._consts is None:
            upper, lower = self.dynamics_model.control_limits
            self._consts = (lower.to(self.device, torch.float32).contiguous(),
                            upper.to(self.device, torch.float32).contiguous())
        return self._consts

    def _filter(self, datax: torch.Tensor, relaxation_penalty: Optional[float] = None,
                u_ref: Optional[torch.Tensor] = None):
        dm = self.dynamics_model
        dev = self.device
        d = datax.to(dev, torch.float32)
        if d.shape[1] != 2 * dm.n_dims + 1:
            raise ValueError(f"datax must be (bs, {2 * dm.n_dims + 1}) = [q | o | do/dq]")
        lo, hi = self._limits()
        K = dm.K_on(dev)
        pen = self.clf_relaxation_penalty if relaxation_penalty is None else relaxation_penalty
        uref = torch.empty(0, device=dev) if u_ref is None else u_ref.to(dev, torch.float32)
        return ops.cbf_filter(d, dm.goal_tensor(dev), self._weights(), K, lo, hi, float(self.clf_lambda), float(pen),
                              self.h_hidden_size, self.h_hidden_layers, uref)

    # ------------------------------------------------------------------ reference API
    def h(self, datax: torch.Tensor):
        """(bs, >= n+1) -> (bs, 1); only the state and the min-distance observation are used (:71-81)."""
        n = self.dynamics_model.n_dims
        if datax.shape[1] < 2 * n + 1:  # h only needs [q, o]; pad the unused do/dq block with zeros
            full = torch.zeros(datax.shape[0], 2 * n + 1, device=datax.device, dtype=datax.dtype)
            full[:, :n + 1] = datax[:, :n + 1]
            datax = full
        h = self._filter(datax)[0]
        return h.reshape(-1, 1).to(datax.device).type_as(datax)

    def h_with_jacobian(self, x: torch.Tensor, data_jacobian: tuple = None):
        """:83-102 -> (h (bs,1), J (bs,1,n), {})"""
        h, J, _, _ = self._filter(x)
        n = self.dynamics_model.n_dims
        return (h.reshape(-1, 1).to(x.device).type_as(x), J.reshape(-1, 1, n).to(x.device).type_as(x), {})

    def V_with_lie_derivatives(self, x: torch.Tensor, data_jacobian: tuple = (), scenarios=None):
        """clf_controller.py:169-218 with f = 0, g = I: Lf_V = 0, Lg_V = J for every scenario."""
        t0 = time.time()
        if scenarios is None:
            scenarios = self.scenarios
        S, n = len(scenarios), self.dynamics_model.n_controls
        h, J, _ = self.h_with_jacobian(x)
        Lf_V = torch.zeros(x.shape[0], S, 1).type_as(x).to(x.device)
        Lg_V = J.expand(-1, S, -1).contiguous()
        t = time.time() - t0
        return h, Lf_V, Lg_V, {"V_w_Jacobian": t, "lie_derivative": 0.0}

    def solve_CLF_QP(self, x, relaxation_penalty: Optional[float] = None, u_ref: Optional[torch.Tensor] = None,
                     requires_grad: bool = False, strict_solution: bool = False):
        """clf_controller.py:461-528 -> ((u (bs,n), r (bs,S)), t_dict).  All scenarios of the arm systems share
        f = 0, g = I, so their constraints coincide and every r_i equals the single-constraint relaxation."""
        t0 = time.time()
        if u_ref is not None:
            assert u_ref.shape[0] == x.shape[0], f"u_ref must have {x.shape[0]} rows, but got {u_ref.shape[0]}"
            assert u_ref.shape[1] == self.dynamics_model.n_controls
        if requires_grad:
            return self._solve_CLF_QP_differentiable(x, relaxation_penalty, u_ref), \
                {"V_w_Jacobian": 0.0, "lie_derivative": 0.0, "nn_forward": 0.0, "QP": time.time() - t0}
        h, J, u, r = self._filter(x, relaxation_penalty, u_ref)
        u = u.to(x.device).type_as(x)
        r = r.reshape(-1, 1).expand(-1, self.n_scenarios).contiguous().to(x.device).type_as(x)
        t = time.time() - t0
        # the four reference timers (clf_controller.py:186-218,487-528); the fused launch is booked on nn_forward
        return (u, r), {"V_w_Jacobian": 0.0, "lie_derivative": 0.0, "nn_forward": t, "QP": 0.0}

    def _solve_CLF_QP_differentiable(self, x, relaxation_penalty, u_ref):
        """The training-time form of the same filter (the reference: autograd through h_nn for h and J, :83-102, then
        backpropagation through the CvxpyLayer, clf_controller.py:354-404): h and J = Jh_q + Jh_o do/dq from torch
        autograd on ``h_nn`` (graph kept, so gradients reach the weights through both), the QP by cbfrrt::qp with its
        closed-form reverse mode cbfrrt::qp_backward (``ops.qp_diff``).  Inference never takes this path."""
        dm, dev = self.dynamics_model, self.device
        n = dm.n_dims
        d = x.to(dev, torch.float32)
        net = self.h_nn.to(dev)
        with torch.enable_grad():
            xin = d[:, :n + 1].detach().requires_grad_(True)
            h = net(xin)
            Jh = torch.autograd.grad(h.sum(), xin, create_graph=True)[0]
            J = Jh[:, :n] + Jh[:, n:] * d[:, n + 1:]
            lo, hi = self._limits()
            ur = dm.u_nominal(d) if u_ref is None else u_ref.to(dev, torch.float32)
            pen = self.clf_relaxation_penalty if relaxation_penalty is None else relaxation_penalty
            u, r = ops.qp_diff(J, float(self.clf_lambda) * h.reshape(-1), ur, lo, hi, float(pen))
        return (u.to(x.device).type_as(x),
                r.reshape(-1, 1).expand(-1, self.n_scenarios).contiguous().to(x.device).type_as(x))

    def u(self, datax: torch.Tensor, u_nominal=None):
        sol, t_dict = self.solve_CLF_QP(datax)
        return sol[0], t_dict

    def forward(self, x):
        return self.u(x)

    def load_weights(self, state_dict):
        """Accepts a Lightning checkpoint's ``state_dict`` (keys h_nn.*) or an h_nn state_dict."""
        sd = {k[len("h_nn."):] if k.startswith("h_nn.") else k: v for k, v in state_dict.items()
              if k.startswith("h_nn.") or k.split(".")[0] in self.h_layers}
        self.h_nn.load_state_dict(sd)
        self._packed = None

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        """Reads a Lightning ``.ckpt`` as the reference's evaluation scripts do
        (motion_planning/evaluation/eval_rrt_mindis_cbfsteer.py:33-50, neural_cbf/evaluation/eval_arm_mindis.py:88):
        a ``torch.save`` dict with ``state_dict`` (keys h_nn.*) and optionally ``hyper_parameters``
        (neural_obs_cbf_controller.py:81-82); keyword arguments override the stored hyper-parameters, as in Lightning.
        ``dynamics_model`` and ``scenarios`` must be passed (they are never stored in the checkpoint)."""
        ckpt = torch.load(checkpoint_path, map_location="cpu" if map_location is None else map_location,
                          weights_only=False)
        if "state_dict" not in ckpt:
            raise KeyError(f"{checkpoint_path}: not a Lightning checkpoint (no 'state_dict')")
        stored = ckpt.get("hyper_parameters", {})
        stored = dict(stored) if isinstance(stored, dict) else dict(vars(stored))     # Lightning dict or a raw Namespace
        hparams = {k: v for k, v in stored.items()
                   if k in ("cbf_alpha", "cbf_relaxation_penalty", "safe_level", "unsafe_level", "cbf_hidden_layers",
                            "cbf_hidden_size", "use_neural_actor", "loss_config", "learn_shape_epochs")}
        hparams.update(kwargs)
        controller = cls(**hparams)
        controller.load_weights(ckpt["state_dict"])
        return controller

