"""Controller -- the abstract base of the controller hierarchy (interface of neural_cbf/controllers/controller.py:7-26:
attributes ``dynamics_model``, ``experiment_suite``, ``controller_period`` and the abstract ``u(x)``)."""
import abc


class Controller(abc.ABC):
    """Holds what every controller of the hierarchy shares; concrete classes implement ``u``."""

    def __init__(self, dynamics_model, experiment_suite=None, controller_period: float = 0.01):
        super().__init__()
        self.dynamics_model, self.experiment_suite = dynamics_model, experiment_suite
        self.controller_period = float(controller_period)

    @abc.abstractmethod
    def u(self, x):
        """control for the batch of states ``x`` -> (bs, n_controls) [, t_dict]"""
        raise NotImplementedError
