"""Planner evaluation over a problem set -- the callers of the whole path in the reference's evaluation scripts:

    motion_planning/evaluation/eval_rrt_all.py:20-71            problem set -> method -> summary + result pickle
    motion_planning/evaluation/eval_batchrrt_vanilla.py         batched RRT, nominal ("line") steer
    motion_planning/evaluation/eval_batchrrt_mindis_cbfsteer.py:19-71   batched RRT, neural-CBF steer from a checkpoint

Same keyword arguments (simulation_dt, controller_period, start_idx, end_idx, RRT_STEP, goal_biasing, max_node,
cbf_alpha, ...), same per-problem protocol (reset_env -> set_goal -> plan with stop_when_success) and the same result
list (one solution dict per problem).  Every steer of a batch is one rollout launch, every nearest-node query one knn
launch, every goal test one masks launch (cbfrrt.motion_planning.baseline.batch_rrt)."""
import os
import pickle
import random

import numpy as np
import torch

from ..baseline import batch_RRT_plan
from .problems import get_test_cases

RESULT_KEYS = ("success", "path", "edge_states", "nominal_path", "path_cost", "explored_nodes", "time_dict")


def seed_everything(seed: int):
    """what pl.seed_everything does for the libraries the planners draw from"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _plan_all(environment, dynamics_model, controller, steer_type, obs_positions, obs_sizes, start_configs, end_configs,
              batch, **kwargs):
    n = min(kwargs.get("end_idx", len(start_configs)) - kwargs.get("start_idx", 0), len(start_configs))
    solutions = []
    for idx in range(n):
        goal_state, init_state = np.asarray(end_configs[idx]), np.asarray(start_configs[idx])
        environment.reset_env(enable_object=False, obs_configs=(obs_positions[idx], obs_sizes[idx]))
        dynamics_model.set_goal(goal_state)
        sol = batch_RRT_plan(env=environment, init_state=init_state, goal_state=goal_state, dynamics_model=dynamics_model,
                             batch=batch, controller=controller, RRT_PARAM=kwargs.get("RRT_STEP", 180),
                             steer_type=steer_type, model_eps=kwargs.get("goal_biasing", 0.2),
                             T=kwargs.get("max_node", 500), stop_when_success=True)
        solutions.append({k: sol[k] for k in RESULT_KEYS})
    return solutions


def eval_batchrrt_vanilla(seed, environment, robot, obs_positions, obs_sizes, start_configs, end_configs, batch=10,
                          **kwargs):
    """eval_batchrrt_vanilla.py:21-47: nominal steer (no network), batches of 10."""
    from ...neural_cbf.systems import ArmDynamics
    seed_everything(seed)
    dynamics_model = ArmDynamics({"name": 1}, dt=kwargs.get("simulation_dt", 1 / 120), dis_threshold=0.02,
                               controller_dt=kwargs.get("controller_period", 1 / 30), env=environment, robot=robot)
    return _plan_all(environment, dynamics_model, None, "line", obs_positions, obs_sizes, start_configs, end_configs,
                     batch, **kwargs)


def eval_batchrrt_mindis(seed, environment, robot, method_model_dir, obs_positions, obs_sizes, start_configs, end_configs,
                         batch=1, controller=None, **kwargs):
    """eval_batchrrt_mindis_cbfsteer.py:19-71: CBF steer with the first ``.ckpt`` of ``method_model_dir`` (or a
    controller built by the caller -- no trained checkpoint ships with the reference)."""
    from ...neural_cbf.controllers import NeuralMindisCBFController
    from ...neural_cbf.systems import ArmMindis
    seed_everything(seed)
    nominal_params = {"name": 1}
    dynamics_model = ArmMindis(nominal_params, dt=kwargs.get("simulation_dt", 1 / 120), dis_threshold=0.02,
                               controller_dt=kwargs.get("controller_period", 1 / 30), env=environment, robot=robot)
    if controller is None:
        model_file = next(os.path.join(method_model_dir, f) for f in sorted(os.listdir(method_model_dir))
                          if f.endswith(".ckpt"))
        controller = NeuralMindisCBFController.load_from_checkpoint(
            model_file, dynamics_model=dynamics_model, scenarios=[nominal_params], datamodule=None,
            experiment_suite=None, cbf_alpha=kwargs.get("cbf_alpha", 0.1), safe_level=0.01, unsafe_level=0.01,
            cbf_hidden_layers=2, cbf_hidden_size=48, use_neural_actor=False, cbf_relaxation_penalty=500)
    else:
        controller.dynamics_model = dynamics_model
    return _plan_all(environment, dynamics_model, controller, "cbf", obs_positions, obs_sizes, start_configs,
                     end_configs, batch, **kwargs)


def summarize(result):
    """the numbers eval_rrt_all.py:55-66 prints"""
    ok = [s for s in result if s["success"]]
    return {"problems": len(result), "success_rate": len(ok) / max(1, len(result)),
            "avg_explored_nodes": float(np.mean([s["explored_nodes"] for s in result])) if result else 0.0,
            "avg_explored_nodes_in_success": float(np.mean([s["explored_nodes"] for s in ok])) if ok else None,
            "total_time": float(np.sum([s["time_dict"]["total_time"] for s in result]))}


METHODS = {"batchrrt_vanilla": eval_batchrrt_vanilla, "batchrrt_mindis": eval_batchrrt_mindis}


def eval_rrt_all(robot_name, method_name, config_file, out_file=None, seed=23, device="cuda", **kwargs):
    """eval_rrt_all.py:20-71: environment from the problem file, cases [start_idx, end_idx), one method, summary,
    result list pickled to ``out_file`` (the reference's ``time_<robot>_<level>_<seed>_...pkl``)."""
    from ...environment import ArmEnv
    environment = ArmEnv([robot_name], GUI=0, config_file=config_file, device=device)
    robot = environment.robot_list[0]
    positions, sizes, start_configs, end_configs = get_test_cases(config_file, kwargs.get("start_idx", 0),
                                                                  kwargs.get("end_idx", None))
    kwargs = dict(kwargs)
    kwargs["start_idx"], kwargs["end_idx"] = 0, len(start_configs)
    method = METHODS[method_name]
    if method is eval_batchrrt_mindis:
        result = method(seed, environment, robot, kwargs.pop("method_model_dir", None), positions, sizes, start_configs,
                        end_configs, **kwargs)
    else:
        result = method(seed, environment, robot, positions, sizes, start_configs, end_configs, **kwargs)
    if out_file:
        with open(out_file, "wb") as handle:
            pickle.dump(result, handle, protocol=pickle.HIGHEST_PROTOCOL)
    return result, summarize(result)
