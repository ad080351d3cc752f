"""NeuralObsCBFController -- mirror of neural_cbf/controllers/neural_obs_cbf_controller.py:28-157,428-480
(inference members).  The Lightning training / validation steps (:159-426) are out of scope; the class is a
plain torch.nn.Module so that ``state_dict()`` keeps the reference's key names (h_nn.*)."""
from typing import Optional, Tuple

import torch

from .cbf_controller import CBFController


class NeuralObsCBFController(torch.nn.Module, CBFController):
    """h(safe) < 0, h(unsafe) > 0, dh/dt <= -lambda h; u from the CBF-QP (neural_obs_cbf_controller.py:29-52)."""

    def __init__(self, dynamics_model, scenarios, datamodule=None, experiment_suite=None, **kwargs):
        torch.nn.Module.__init__(self)
        CBFController.__init__(self, dynamics_model=dynamics_model, scenarios=scenarios,
                               experiment_suite=experiment_suite, cbf_lambda=kwargs["cbf_alpha"],
                               cbf_relaxation_penalty=kwargs["cbf_relaxation_penalty"],
                               controller_period=dynamics_model.controller_dt,
                               safe_level=kwargs.get("safe_level", 0.0), unsafe_level=kwargs.get("unsafe_level", 0.0))
        self.dynamics_model = dynamics_model
        self.datamodule = datamodule
        self.cbf_alpha = kwargs["cbf_alpha"]
        assert self.cbf_alpha > 0
        assert self.cbf_alpha * self.dynamics_model.controller_dt <= 1
        self.use_neural_actor = kwargs.get("use_neural_actor", False)
        self.h_hidden_layers = kwargs["cbf_hidden_layers"]
        self.h_hidden_size = kwargs["cbf_hidden_size"]
        if "loss_config" in kwargs:
            self.loss_config = kwargs["loss_config"]
            self.learn_shape_epochs = kwargs.get("learn_shape_epochs")

    def h(self, datax: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def h_with_jacobian(self, datax: torch.Tensor, data_jacobian: tuple) -> Tuple[torch.Tensor, torch.Tensor, dict]:
        raise NotImplementedError

    def V(self, x: torch.Tensor):
        return self.h(x)

    def V_with_jacobian(self, x: torch.Tensor, data_jacobian: tuple = ()):
        return self.h_with_jacobian(x, data_jacobian)

    def forward(self, x):
        assert x.shape[1] == self.n_dims_extended
        return self.u(x)

    def u(self, datax: torch.Tensor, u_nominal=None):
        """neural_obs_cbf_controller.py:428-480: the neural-actor branch is dead code (`if False`)."""
        return CBFController.u(self, datax)
