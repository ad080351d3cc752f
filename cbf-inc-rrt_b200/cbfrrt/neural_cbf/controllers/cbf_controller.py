"""CBFController -- mirror of neural_cbf/controllers/cbf_controller.py:11-67 (h = V - 1 re-using the CLF QP)."""
import torch

from .clf_controller import CLFController


class CBFController(CLFController):
    def __init__(self, dynamics_model, scenarios, experiment_suite=None, cbf_lambda: float = 1.0,
                 cbf_relaxation_penalty: float = 5000.0, controller_period: float = 0.01, **kwargs):
        super(CBFController, self).__init__(dynamics_model=dynamics_model, scenarios=scenarios,
                                            experiment_suite=experiment_suite, clf_lambda=cbf_lambda,
                                            clf_relaxation_penalty=cbf_relaxation_penalty,
                                            controller_period=controller_period, **kwargs)

    def V_with_jacobian(self, x: torch.Tensor, data_jacobian: tuple = ()):
        V, JV, d = super().V_with_jacobian(x, data_jacobian)
        return V - 1.0, JV, d
