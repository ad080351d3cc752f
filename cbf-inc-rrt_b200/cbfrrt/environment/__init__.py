"""Mirror of the reference's ``environment`` package (environment/__init__.py:1-9)."""
from .arm_env import ArmEnv
from .basic_robot import BasicRobot
from .franka_panda import FrankaPanda
from .magician import Magician

__all__ = ["ArmEnv", "BasicRobot", "FrankaPanda", "Magician"]
