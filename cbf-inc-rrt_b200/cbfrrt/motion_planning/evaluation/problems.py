"""Planning-problem sets in the reference's on-disk formats.

    motion_planning/evaluation/generate_problem_env.py:19-92   proposal + filtering of problems, raw ``pkl/`` stream
    motion_planning/evaluation/collect_problems.py:9-33        raw stream -> ``refined_pkl/<robot>_<level>.pkl``
    motion_planning/evaluation/eval_rrt_all.py:74-104          ``get_test_cases`` (reads the refined file)
    environment/arm_env.py:55,103-111                          ``ArmEnv(config_file=...)`` reads the same file

Raw stream: records appended with ``pickle.dump`` ('ab'), each the planner's solution dict plus
``obstacle_positions (O,3)``, ``obstacle_sizes (O,3)``, ``init_config (n,)``, ``goal_config (n,)``.
Refined file: ONE pickled dict with the same keys, every value a list over problems.  ``np.load(...,
allow_pickle=True)`` opens both a refined pickle and an ``.npz`` with those keys, which is how the reference (and
``ArmEnv`` here) reads them.

The proposal loop draws from ``np.random`` in the reference's order; its ``sample_safe`` is the batched rejection
sampler of the dynamics model (one masks launch per candidate batch instead of a pybullet loop per state).
"""
import pickle

import numpy as np

PROBLEM_KEYS = ("obstacle_positions", "obstacle_sizes", "init_config", "goal_config")


def propose_test_cases(dynamics_model, obstacle_num):
    """generate_problem_env.py:59-92 -> (positions (O,3), sizes (O,3), start (n,), goal (n,))"""
    while True:
        try:
            sizes = np.zeros((0, 3))
            for ii in range(obstacle_num):
                idx = np.random.randint(0, 3)
                size = np.random.uniform(low=(0.05, 0.05, 0.05), high=(0.25, 0.25, 0.25), size=(1, 3))
                if idx == 0:    # pad
                    size[0, ii % 3] = 0.05
                elif idx == 1:  # stick
                    size[0, ii % 3] = 0.05
                    size[0, (ii + 1) % 3] = 0.05
                sizes = np.concatenate((sizes, size), axis=0)
            while True:
                positions = np.random.uniform(low=(-0.5, -0.5, 0), high=(0.5, 0.5, 1), size=(obstacle_num, 3))
                low = positions[:, 2] - sizes[:, 2] < 0.3
                near = (np.abs(np.abs(positions[:, 0]) - sizes[:, 0]) < 0.1) | \
                       (np.abs(np.abs(positions[:, 1]) - sizes[:, 1]) < 0.1)
                if not np.any(low & near):   # every obstacle "far from the base" (:79-81)
                    break
            dynamics_model.env.reset_env(obs_configs=(positions, sizes), enable_object=False)
            for _ in range(10):
                safe_configs = dynamics_model.sample_safe(2)
                start_config = safe_configs[0, :dynamics_model.n_dims].cpu().numpy()
                end_config = safe_configs[1, :dynamics_model.n_dims].cpu().numpy()
                if np.linalg.norm(start_config - end_config) > 0.7:
                    return positions, sizes, start_config, end_config
        except RuntimeWarning:
            continue


def append_problem(file_name, solution, positions, sizes, start_config, end_config):
    """generate_problem_env.py:51-56: one record appended to the raw stream."""
    record = dict(solution)
    record.update({"obstacle_positions": positions, "obstacle_sizes": sizes, "init_config": start_config,
                   "goal_config": end_config})
    with open(file_name, "ab") as handle:
        pickle.dump(record, handle, protocol=pickle.HIGHEST_PROTOCOL)
    return record


def generate_refined_test_cases(dynamics_model, solver, file_name, problem_num=200, obstacle_num=8, seed=1234,
                                max_proposals=None):
    """generate_problem_env.py:19-56 with the planner passed in: ``solver(env, init_state, goal_state, dynamics_model)``
    returns the planner's solution dict (the reference uses BIT* with steer_type='line', batch_size=50); a proposal is
    kept when ``solution['success']`` and ``solution['time_dict']['total_time']`` are truthy.  Seeds ``np.random`` /
    ``torch`` / ``random`` like ``pl.seed_everything`` (:35-36).  -> number of problems written."""
    import random

    import torch
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    env = dynamics_model.env
    written = proposals = 0
    while written < problem_num and (max_proposals is None or proposals < max_proposals):
        proposals += 1
        positions, sizes, start_config, end_config = propose_test_cases(dynamics_model, obstacle_num=obstacle_num)
        env.reset_env(enable_object=False, obs_configs=(positions, sizes))
        dynamics_model.set_goal(end_config)
        solution = solver(env, start_config, end_config, dynamics_model)
        if solution["success"] and solution["time_dict"]["total_time"]:
            append_problem(file_name, solution, positions, sizes, start_config, end_config)
            written += 1
    return written


def read_problem_stream(file_name):
    """collect_problems.py:12-18: every record of a raw stream, in order."""
    records = []
    with open(file_name, "rb") as handle:
        while True:
            try:
                records.append(pickle.load(handle))
            except EOFError:
                break
    return records


def collect_problems(records, t_min=0.5, t_max=2.0, limit=1000):
    """collect_problems.py:20-30: dict of lists over the records whose planning time lies in (t_min, t_max)."""
    if not records:
        return {}
    collected = {key: [] for key in records[0].keys()}
    kept = 0
    for rec in records:
        if t_min < rec["time_dict"]["total_time"] < t_max:
            for key in collected:
                collected[key].append(rec[key])
            kept += 1
        if kept >= limit:
            break
    return collected


def save_collected(file_name, collected):
    """collect_problems.py:32-33"""
    with open(file_name, "wb") as handle:
        pickle.dump(collected, handle, protocol=pickle.HIGHEST_PROTOCOL)


def get_test_cases(file_name, start_idx=0, end_idx=None, **kwargs):
    """eval_rrt_all.py:74-104 -> (positions, sizes, start_configs, end_configs) of problems [start_idx, end_idx);
    accepts the singular (``init_config``) and plural (``init_configs``) key spellings."""
    import os
    if not os.path.exists(file_name):
        raise ValueError("No test cases found!")
    data = np.load(file_name, allow_pickle=True)
    sl = slice(start_idx, end_idx)
    positions = data["obstacle_positions"][sl]
    sizes = data["obstacle_sizes"][sl]
    try:
        start_configs, end_configs = data["init_config"][sl], data["goal_config"][sl]
    except KeyError:
        start_configs, end_configs = data["init_configs"][sl], data["goal_configs"][sl]
    return positions, sizes, start_configs, end_configs
