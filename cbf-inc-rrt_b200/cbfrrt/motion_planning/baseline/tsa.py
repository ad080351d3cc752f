"""steer / edge_checking -- mirror of motion_planning/baseline/tsa.py:217-286.

  * ``steer``: the reference's host loop verbatim (same signature, same return values), running on this package's
    dynamics_model / controller, i.e. every call inside the loop is a CUDA op;
  * ``steer_fused``: the whole loop in ONE launch (cbfrrt::rollout) with identical semantics;
  * ``edge_checking``: the reference's signature and result; its K+1 safe_mask calls are one launch
    (``edge_checking_reference_loop`` keeps the per-point loop for parity tests);
  * ``edge_checking_eps`` / ``edge_checking_batch``: many edges per launch with per-edge K = int(d/eps)
    (cbfrrt::edge_validity_var) or one K for all (cbfrrt::edge_validity) -- what the planners should call.
Tree growth, nearest-neighbour search and BIT* queues stay host Python.
"""
import time

import numpy as np
import torch

from ... import ops


@torch.no_grad()
def steer(u, dynamics_model, new_state, nearest, RRT_step=10, device="cpu"):
    """tsa.py:217-269, unchanged control flow -> (x_last, no_collision, time_dict, x_list)."""
    time_dict = {"complete_sample": 0, "cl_dynamics": 0, "misc": 0, "collision_checking": 0}
    t0 = time.time()
    no_collision = True
    x_list = [nearest]
    controller_update_freq = dynamics_model.controller_dt // dynamics_model.dt
    dynamics_model.set_intermediate_goals(new_state)
    x_current = dynamics_model.complete_sample_with_observations(
        torch.tensor(nearest, device=device, dtype=torch.float32).unsqueeze(0), 1)
    time_dict["complete_sample"] += (time.time() - t0)
    for tstep in range(RRT_step):
        if tstep % controller_update_freq == 0:
            tt0 = time.time()
            u_current = u(x_current)
            if isinstance(u_current, tuple):
                u_current, u_time_dict = u_current
            else:
                u_time_dict = {"u": time.time() - tt0}
            for key in u_time_dict.keys():
                time_dict[key] = time_dict.get(key, 0) + u_time_dict[key]
        x_current, tt_list = dynamics_model.closed_loop_dynamics(
            x_current, u_current, update_observation=((tstep + 1) % controller_update_freq == 0), return_time=True)
        time_dict["cl_dynamics"] += tt_list[0]
        time_dict["complete_sample"] += tt_list[1]
        tt1 = time.time()
        if dynamics_model.unsafe_mask(x_current):
            no_collision = False
            break
        tt2 = time.time()
        x_list.append(x_current.cpu().squeeze().numpy()[:dynamics_model.q_dims])
        if tstep > 10 and np.linalg.norm(x_list[-1] - x_list[-10]) < 1e-1:  # stuck
            break
        time_dict["collision_checking"] += (tt2 - tt1)
        time_dict["misc"] += (time.time() - tt2)
    return x_list[-1], no_collision, time_dict, x_list


@torch.no_grad()
def steer_fused(controller, dynamics_model, new_state, nearest, RRT_step=10, device="cpu"):
    """Same contract as ``steer(controller.u, ...)`` in one kernel launch (cbfrrt::rollout)."""
    t0 = time.time()
    dm = dynamics_model
    dm.set_intermediate_goals(np.asarray(new_state))
    q0 = torch.as_tensor(np.asarray(nearest), dtype=torch.float32).reshape(1, -1).to(dm.device)
    qf, alive, done, traj = _rollout(controller, dm, q0, RRT_step, stuck=(9, 10, 0.1), want_traj=True)
    k = int(done.item())
    x_list = [nearest] + [row for row in traj[0, :k].cpu().numpy()]
    time_dict = {"complete_sample": 0, "cl_dynamics": 0, "misc": 0, "collision_checking": 0,
                 "fused_rollout": time.time() - t0}
    return x_list[-1], bool(alive.item()), time_dict, x_list


def _rollout(controller, dm, q0, steps, stuck=(0, 10, 0.1), want_traj=False, keep_going=False, stuck_mode=0):
    b, s, fl, gr = dm.env.scene
    lo, hi = controller._limits()
    ctrl_every = int(dm.controller_dt // dm.dt)
    return ops.rollout(q0, dm.goal_tensor(dm.device), int(steps), ctrl_every, float(dm.dt), int(stuck[0]),
                       int(stuck[1]), float(stuck[2]), bool(want_traj), bool(keep_going), b, s, fl, gr, controller._weights(),
                       dm.K_on(dm.device), lo, hi, float(controller.clf_lambda),
                       float(controller.clf_relaxation_penalty), float(dm.dis_threshold), float(dm.thr_hard),
                       dm.robot.robot_id, controller.h_hidden_size, controller.h_hidden_layers, int(stuck_mode))


def edge_checking(state, new_state, dynamics_model, eps):
    """tsa.py:272-286: safe(new_state) and safe(state + k/K (new_state - state)), k = 0..K-1, K = int(d / eps) -- the
    reference's K+1 single-state safe_mask calls collapse into ONE launch over the edge's points."""
    assert state.size == new_state.size
    valid, _ = edge_checking_eps([state], [new_state], dynamics_model, eps)
    return bool(valid[0])


def edge_checking_reference_loop(state, new_state, dynamics_model, eps):
    """the reference's loop verbatim in structure (one safe_mask call per point, early exit); kept for parity tests"""
    assert state.size == new_state.size
    if not dynamics_model.safe_mask(torch.Tensor(new_state).unsqueeze(0)):
        return False
    disp = new_state - state
    d = np.linalg.norm(state - new_state)
    K = int(d / eps)
    for k in range(0, K):
        c = state + k * 1. / K * disp
        if not dynamics_model.safe_mask(torch.Tensor(c).unsqueeze(0)):
            return False
    return True


def edge_checking_eps(states, new_states, dynamics_model, eps):
    """E edges, each with its own K_e = int(|new - old| / eps) (tsa.py:277-279, computed in float64 on the host like
    the reference), in one launch (cbfrrt::edge_validity_var) -> (valid bool[E], first_bad i32[E])."""
    dm = dynamics_model
    env = dm.env
    a64 = np.asarray(states, dtype=np.float64).reshape(-1, dm.q_dims)
    b64 = np.asarray(new_states, dtype=np.float64).reshape(-1, dm.q_dims)
    K = (np.linalg.norm(a64 - b64, axis=1) / eps).astype(np.int64)
    # planners check a handful of edges per call: numpy in / numpy out (cbfrrt_edge_validity_host), one copy each way
    d = torch.device(dm.device)
    valid, first_bad = ops.edge_validity_host(dm.robot.name, a64, b64, K, np.concatenate([env.obs_positions, env.obs_sizes], 1),
                                              env.obs_spheres, bool(env.include_floor), float(dm.dis_threshold), float(dm.thr_hard),
                                              device=(d.index if d.type == "cuda" and d.index is not None else 0))
    return torch.from_numpy(valid), torch.from_numpy(first_bad)


def edge_checking_batch(states, new_states, dynamics_model, n_interp):
    """E edges x (n_interp + 1) points in one launch (cbfrrt::edge_validity) -> (valid bool[E], first_bad i32[E])."""
    dm = dynamics_model
    b, s, fl, gr = dm.env.scene
    qf = torch.as_tensor(np.asarray(states), dtype=torch.float32).reshape(-1, dm.q_dims).to(dm.device)
    qt = torch.as_tensor(np.asarray(new_states), dtype=torch.float32).reshape(-1, dm.q_dims).to(dm.device)
    valid, first_bad = ops.edge_validity(qf, qt, int(n_interp), b, s, fl, gr, float(dm.dis_threshold), float(dm.thr_hard),
                                         dm.robot.robot_id)
    return valid.bool(), first_bad
