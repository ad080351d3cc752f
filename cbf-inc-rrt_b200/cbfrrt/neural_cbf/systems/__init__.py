"""Mirror of neural_cbf/systems/__init__.py (ArmLidar is a "next" row)."""
from .control_affine_system import ControlAffineSystem
from .arm_dynamics import ArmDynamics
from .arm_mindis import ArmMindis

__all__ = ["ControlAffineSystem", "ArmDynamics", "ArmMindis"]
