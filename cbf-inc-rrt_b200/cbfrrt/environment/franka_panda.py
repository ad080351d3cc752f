"""FrankaPanda -- mirror of environment/franka_panda.py:17-31."""
import numpy as np

from .basic_robot import BasicRobot


class FrankaPanda(BasicRobot):
    def __init__(self, bullet_client=None, device="cuda"):
        super().__init__("panda", device=device)
        self.q0 = np.array(self.model["q0"])  # franka_panda.py:23-24, body-only
        self.invalid_link = [7, 11]            # franka_panda.py:25
        self.set_joint_position(self.body_joints, self.q0)

    def __str__(self):
        return "Franka_Panda"
