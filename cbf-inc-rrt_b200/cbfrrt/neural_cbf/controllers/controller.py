"""Controller -- mirror of neural_cbf/controllers/controller.py:7-26."""
from abc import ABC, abstractmethod


class Controller(ABC):
    controller_period: float

    def __init__(self, dynamics_model, experiment_suite=None, controller_period: float = 0.01):
        super(Controller, self).__init__()
        self.controller_period = controller_period
        self.dynamics_model = dynamics_model
        self.experiment_suite = experiment_suite

    @abstractmethod
    def u(self, x):
        """Get the control input for a given state"""
