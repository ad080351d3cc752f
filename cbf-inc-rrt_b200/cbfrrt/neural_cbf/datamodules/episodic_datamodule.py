"""EpisodicDataModule -- mirror of neural_cbf/datamodules/episodic_datamodule.py:21-460 without Lightning.

The data module is the bulk caller of the sampled-state sweeps: ``sample_fixed`` (:146-194) draws the quota of safe /
unsafe / boundary / goal states per training scene and labels them with the four masks; ``sample_trajectories``
(:91-144) rolls the noisy nominal controller out from boundary states.  Here every one of those calls is a batched
launch on the dynamics model (rejection sampling over whole candidate batches, one masks launch per label set)
instead of a pybullet loop per state.  The on-disk format is the reference's (:249-260): one ``torch.save`` dict

    {'x_training', 'x_validation': (N, 2n+1) float32 [q | o | do/dq],
     'x_training_mask', 'x_validation_mask': {'safe','unsafe','goal','boundary': (N,) bool}}

under ``<root>/models/neural_cbf/<str(model)>/data/<thr>_<obstacles>_<noise>_<episodes>_<traj>_<len>_<fixed>.pt``
(:203-218), so datasets written by either side load on the other.  Host-side RNG calls (``random``, ``torch``) are
made in the reference's order: with the same seeds the same candidates are drawn.

Not here: the LiDAR branch (:66-67,205-212,261-267,283-305) and motor-control rollouts (the
reference steps the physics engine for 80 % of the episodes, :120; this path is kinematic, see ``motor_control``).
"""
import os
import random
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch.utils.data import DataLoader, TensorDataset

MASK_KEYS = ("safe", "unsafe", "goal", "boundary")


class EpisodicDataModule:
    def __init__(self, model, initial_domain: List[Tuple[float, float]], max_episode: int = 1,
                 trajectories_per_episode: int = 100, trajectory_length: int = 5000, fixed_samples: int = 100000,
                 val_split: float = 0.1, batch_size: int = 64, noise_level: float = 0, total_point: int = 0,
                 quotas: Optional[Dict[str, float]] = None, shuffle: bool = True, name: str = "",
                 data_root: Optional[str] = None, num_workers: int = 0):
        self.name = name
        self.model = model
        self.n_dims = model.n_dims
        self.shuffle = shuffle
        if "lidar" in str(self.model):
            raise NotImplementedError("LiDAR observation models are outside the hot path (SURVEY 8f rank 2)")
        self.max_episode = max_episode
        self.trajectories_per_episode = trajectories_per_episode
        self.trajectory_length = trajectory_length
        self.fixed_samples = fixed_samples
        self.val_split = val_split
        self.batch_size = batch_size
        self.noise_level = noise_level
        self.quotas = quotas if quotas is not None else {}
        assert len(initial_domain) == self.n_dims
        self.initial_domain = initial_domain
        self.x_max, self.x_min = model.state_limits
        self.x_center = (self.x_max + self.x_min) / 2.0
        self.x_range = self.x_max - self.x_min
        self.data_root = os.getcwd() if data_root is None else data_root
        self.num_workers = num_workers

    # ------------------------------------------------------------------ labels
    def _label(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The four masks of :135-138 / :182-185 from ONE masks launch (the reference runs the two loops 5 times)."""
        m = self.model
        if hasattr(m, "masks"):
            safe, unsafe = m.masks(x)
            goal = (x[:, :m.q_dims] - torch.as_tensor(m.goal_state[:m.q_dims]).type_as(x)).norm(dim=1) <= 0.2
            return {"safe": safe, "unsafe": unsafe, "goal": goal & safe, "boundary": ~(safe | unsafe)}
        return {"safe": m.safe_mask(x), "unsafe": m.unsafe_mask(x), "goal": m.goal_mask(x),
                "boundary": m.boundary_mask(x)}

    # ------------------------------------------------------------------ :91-144
    def sample_trajectories(self, simulator: Callable, episode_num=0, trajectory_num=0, motor_control: bool = False):
        if motor_control:
            raise NotImplementedError("motor control needs a physics step (arm_dynamics.py:499-506); out of scope")
        x_sim = []
        x_mask = {k: [] for k in MASK_KEYS}
        for _ in range(episode_num if episode_num else self.max_episode):
            complete_episode = False
            while not complete_episode:
                try:
                    self.model.env.reset_env(
                        obs_configs=self.model.env.get_env_config(int(random.uniform(0, 600))), enable_object=False)
                    self.model.set_goal(self.model.sample_state_space(1)[0, :self.model.q_dims])
                    traj_number = trajectory_num if trajectory_num else self.trajectories_per_episode
                    x_init = self.model.sample_boundary(traj_number, max_tries=2)[:, :self.n_dims]
                    complete_episode = True
                except RuntimeWarning:
                    pass
            random.random()  # the reference's motor-control coin (:120), drawn to keep the host RNG stream aligned
            batch_x_sim = simulator(x_init, self.trajectory_length, noise_level=self.noise_level,
                                    collect_dataset=True, use_motor_control=False)
            if hasattr(self.model, "o_dims"):
                batch_x_sim = batch_x_sim.reshape(-1, self.n_dims + self.model.o_dims + self.model.state_aux_dims)
            else:
                batch_x_sim = batch_x_sim.reshape(-1, self.n_dims)
            x_sim.append(batch_x_sim)
            for k, v in self._label(batch_x_sim).items():
                x_mask[k].append(v)
        return torch.cat(x_sim, dim=0), {k: torch.cat(v, dim=0) for k, v in x_mask.items()}

    # ------------------------------------------------------------------ :146-194
    def sample_fixed(self):
        samples = []
        x_mask = {k: [] for k in MASK_KEYS}
        add_env_idx = 0
        for episode in range(self.max_episode):
            complete_episode = False
            while not complete_episode:
                try:
                    sample = []
                    self.model.env.reset_env(obs_configs=self.model.env.get_env_config(episode + add_env_idx),
                                             enable_object=False)
                    for region_name, quota in self.quotas.items():
                        num_samples = int(self.fixed_samples * quota)
                        if region_name == "goal":
                            sample.append(self.model.sample_goal(num_samples))
                        elif region_name == "safe":
                            sample.append(self.model.sample_safe(num_samples))
                        elif region_name == "unsafe":
                            sample.append(self.model.sample_unsafe(num_samples))
                        elif region_name == "boundary":
                            sample.append(self.model.sample_boundary(num_samples))
                    complete_episode = True
                except RuntimeWarning:
                    add_env_idx += 1
                    continue
            sample = torch.vstack(sample)
            samples.append(sample)
            for k, v in self._label(sample).items():
                x_mask[k].append(v)
        return torch.vstack(samples), {k: torch.cat(v, dim=0) for k, v in x_mask.items()}

    # ------------------------------------------------------------------ :196-330
    def dataset_path(self) -> str:
        """:203-218 (mindis naming)."""
        if "mindis" not in str(self.model):
            raise NotImplementedError(f"Unknown model {self.model}")
        d = os.path.join(self.data_root, "models", "neural_cbf", str(self.model), "data")
        return os.path.join(d, f"{int(100 * self.model.dis_threshold)}_{self.model.env.obstacle_num}_"
                               f"{int(self.noise_level * 1e2)}_{self.max_episode}_{self.trajectories_per_episode}_"
                               f"{self.trajectory_length}_{self.fixed_samples}.pt")

    def prepare_data(self, verbose: bool = False):
        path = self.dataset_path()
        os.makedirs(os.path.dirname(path), exist_ok=True)
        if os.path.exists(path):
            data = torch.load(path)
            self.x_training = data["x_training"]
            self.x_validation = data["x_validation"]
            self.x_training_mask = data["x_training_mask"]
            self.x_validation_mask = data["x_validation_mask"]
        else:
            x_sample, x_sample_mask = self.sample_fixed()
            x_sim, x_sim_mask = self.sample_trajectories(self.model.noisy_simulator)
            x = torch.cat((x_sim, x_sample), dim=0)
            x_mask = {k: torch.cat((x_sim_mask[k], x_sample_mask[k]), dim=0) for k in x_sim_mask.keys()}
            random_indices = torch.randperm(x.shape[0])
            val_pts = int(x.shape[0] * self.val_split)
            validation_indices = random_indices[:val_pts]
            training_indices = random_indices[val_pts:]
            self.x_training = x[training_indices]
            self.x_validation = x[validation_indices]
            self.x_training_mask = {k: x_mask[k][training_indices] for k in x_mask.keys()}
            self.x_validation_mask = {k: x_mask[k][validation_indices] for k in x_mask.keys()}
            torch.save({"x_training": self.x_training, "x_validation": self.x_validation,
                        "x_training_mask": self.x_training_mask, "x_validation_mask": self.x_validation_mask}, path)
        if verbose:
            print(f"Full dataset: {self.x_training.shape[0]} training, {self.x_validation.shape[0]} validation")
            for k in ("goal", "safe", "unsafe", "boundary"):
                print(f"\t{self.x_training_mask[k].sum()} {k} points ({self.x_validation_mask[k].sum()} val)")
        # :306-319: mindis datasets carry (x, goal, safe, unsafe)
        self.training_data = TensorDataset(self.x_training, self.x_training_mask["goal"],
                                           self.x_training_mask["safe"], self.x_training_mask["unsafe"])
        self.validation_data = TensorDataset(self.x_validation, self.x_validation_mask["goal"],
                                             self.x_validation_mask["safe"], self.x_validation_mask["unsafe"])

    def setup(self, stage=None):
        pass

    def train_dataloader(self):
        return DataLoader(self.training_data, batch_size=self.batch_size, num_workers=self.num_workers, shuffle=True)

    def val_dataloader(self):
        return DataLoader(self.validation_data, batch_size=self.batch_size, num_workers=self.num_workers,
                          shuffle=True)
