"""Magician -- mirror of environment/magician.py:18-25 (URDF loaded with globalScaling=2)."""
import numpy as np

from .basic_robot import BasicRobot


class Magician(BasicRobot):
    def __init__(self, bullet_client=None, device="cuda"):
        super().__init__("magician", device=device)
        self.q0 = np.array(self.model["q0"])  # magician.py:22, body-only

    def __str__(self):
        return "Magician"
