"""Mirror of neural_cbf/controllers/__init__.py (inference classes on the hot path)."""
from .controller import Controller
from .clf_controller import CLFController
from .cbf_controller import CBFController
from .neural_obs_cbf_controller import NeuralObsCBFController
from .neural_mindis_cbf_controller import NeuralMindisCBFController
from .neural_lidar_cbf_controller import NeuralLidarCBFController, LidarLayout

__all__ = ["Controller", "CLFController", "CBFController", "NeuralObsCBFController", "NeuralMindisCBFController",
           "NeuralLidarCBFController", "LidarLayout"]
