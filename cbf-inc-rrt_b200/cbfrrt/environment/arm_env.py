"""ArmEnv -- mirror of environment/arm_env.py:23-182 without pybullet.

The scene (optional floor plane z=0 + axis-aligned boxes, arm_env.py:63-99,154-182) is kept on the host in the
reference's own attributes (obs_positions, obs_sizes, obstacle_ids) and mirrored on the GPU in the library's
packed layout; ``reset_env`` re-uploads it.  Sphere obstacles are an extension (BASELINE.json config 3).
"""
import os

import numpy as np
import torch

from .. import ops
from .franka_panda import FrankaPanda
from .magician import Magician


class _NoBullet:
    """Stands where the reference keeps its BulletClient (``env.p``).  Nothing on the B200 path calls it."""

    def __getattr__(self, name):
        raise AttributeError(f"env.p.{name}: pybullet is not part of the B200 path; the scene lives in "
                             f"ArmEnv.obs_positions / obs_sizes and on the GPU (ArmEnv.scene)")


def sample_surface(center, half_extent, idx, num, add_normal=False):
    """arm_env.py:12-20: ``num`` uniform samples on face ``idx`` (0..5) of an axis-aligned box."""
    a = np.zeros((num, 3))
    b = np.random.uniform(low=(-1, -1, -1), high=(1, 1, 1), size=(num, 3))
    for ii in range(3):
        a[:, ii] = center[ii] + half_extent[ii] * b[:, ii]
    a[:, idx // 2] = center[idx // 2] + half_extent[idx // 2] * np.sign(idx % 2 - 0.5)
    if add_normal:
        nrm = np.sign(idx % 2 - 0.5) * np.eye(3)[idx // 2].reshape(1, 3).repeat(num, axis=0)
        a = np.concatenate((a, nrm), axis=1)
    return a


class ArmEnv:
    def __init__(self, robot_name_list, config_file: str = "", GUI=False, include_floor=True, device="cuda"):
        assert not GUI, "no GUI on the B200 path"
        self.robot_name_list = robot_name_list
        self.robot_list = []
        self.obstacle_ids = []
        self.include_floor = include_floor
        self.device = torch.device(device)
        self.p = _NoBullet()
        self.use_training_env = (len(config_file) == 0)
        if self.use_training_env:
            # arm_env.py:44-53: the training file models/env_file/env_600_4.npz is generated on first use
            self.obstacle_num = 4
            self.env_config_file = ""
            self.env_config = self._generate_env_config(None, 4)
        else:
            self.env_config_file = config_file
            self.env_config = dict(np.load(self.env_config_file, allow_pickle=True))
        self.obs_spheres = np.zeros((0, 4))
        self.reset_env(self.get_env_config(-1), enable_object=False)

    def __str__(self):
        return str(self.robot_name_list)

    def reset_env(self, obs_configs, enable_object=False, spheres=None):
        """arm_env.py:63-99.  ``spheres`` (P,4) [cx,cy,cz,r] is an extension of the reference scene."""
        self.robot_list = []
        assert len(self.robot_name_list) <= 1, "Only support one robot now."
        for robot_name in self.robot_name_list:
            if robot_name == "panda":
                self.robot_list.append(FrankaPanda(device=self.device))
            elif robot_name == "magician":
                self.robot_list.append(Magician(device=self.device))
            else:
                raise NotImplementedError(f"Robot {robot_name} not supported yet.")
        if enable_object:
            raise NotImplementedError("grasped objects are outside the hot path (arm_env.py:86-97)")
        self.obs_spheres = np.zeros((0, 4)) if spheres is None else np.asarray(spheres, dtype=np.float64).reshape(-1, 4)
        self._generate_obstacle(obs_configs=obs_configs)

    def get_env_config(self, idx):
        """arm_env.py:103-114"""
        if idx >= 0:
            obs_positions = self.env_config["obstacle_positions"][idx]
            if "obstacle_sizes" in self.env_config.keys():
                obs_sizes = self.env_config["obstacle_sizes"][idx]
            else:
                obs_sizes = np.array([[0.05, 0.05, 0.1] for _ in range(obs_positions.shape[0])])
            return obs_positions, obs_sizes
        return np.array([[0.3, 0.17, 0.3], [-0.4, 0.23, 0.4], [0.0, -0.23, 0.6]]), \
            np.array([[0.05, 0.05, 0.1], [0.05, 0.05, 0.1], [0.05, 0.05, 0.1]])

    def sample_obstacle_surface(self, total_num, add_normal=False):
        """arm_env.py:116-135"""
        n_obs = len(self.obstacle_ids) - int(self.include_floor)
        each_obstacle = np.random.randint(low=0, high=n_obs, size=total_num)
        points_global = np.zeros((0, 3 + 3 * int(add_normal)))
        for obs_idx in range(n_obs):
            rand_idx = np.random.randint(0, 6, np.where(each_obstacle == obs_idx)[0].shape[0])
            for jj in range(6):
                new_array = sample_surface(self.obs_positions[obs_idx], self.obs_sizes[obs_idx], jj,
                                           np.where(rand_idx == jj)[0].shape[0], add_normal=add_normal)
                points_global = np.concatenate((points_global, new_array), axis=0)
        return points_global

    # ===================== internal module ===========================
    def _generate_env_config(self, file_name, obstacle_num, problem_num=1000):
        """arm_env.py:139-152 (kept in memory; written only when a file name is given)."""
        positions = np.random.uniform(low=(-0.5, -0.5, 0), high=(0.5, 0.5, 1), size=(problem_num, obstacle_num, 3))
        if file_name:
            assert file_name.endswith(".npz")
            os.makedirs(os.path.dirname(file_name) or ".", exist_ok=True)
            np.savez(file_name, obstacle_positions=positions)
        return {"obstacle_positions": positions}

    def _generate_obstacle(self, obs_configs):
        """arm_env.py:154-164: obstacle_ids = [plane] + one id per box; then upload the packed scene."""
        self.obs_positions, self.obs_sizes = obs_configs
        self.obs_positions = np.asarray(self.obs_positions, dtype=np.float64).reshape(-1, 3)
        self.obs_sizes = np.asarray(self.obs_sizes, dtype=np.float64).reshape(-1, 3)
        self.plane = 0
        self.obstacle_ids = [self.plane] if self.include_floor else []
        first = 1  # body ids: 0 = plane (or unused), boxes and spheres follow
        self.obstacle_ids += list(range(first, first + self.obs_positions.shape[0] + self.obs_spheres.shape[0]))
        self.scene_boxes, self.scene_spheres = ops.pack_scene(self.obs_positions, self.obs_sizes, self.obs_spheres,
                                                              device=self.device)
        # exact broad phase, rebuilt with the scene (one small kernel); CPU tensors (tests with fake ops) get none
        n_obs = self.scene_boxes.shape[0] + self.scene_spheres.shape[0]
        self.scene_grid = (ops.build_grid(self.scene_boxes, self.scene_spheres)
                           if (self.device.type == "cuda" and n_obs > 0) else ops.no_grid(self.device))

    @property
    def scene(self):
        """(boxes (O,8), spheres (P,4), floor, grid) on the GPU, as the ops take them."""
        return self.scene_boxes, self.scene_spheres, int(self.include_floor), self.scene_grid
