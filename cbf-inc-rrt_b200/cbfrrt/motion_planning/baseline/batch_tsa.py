"""Batched steer and the nearest-neighbour front-end -- mirror of motion_planning/baseline/batch_tsa.py:129-145,191-231.

``steer`` keeps the reference's host loop (per-step batched CUDA ops replace its per-row pybullet loops,
arm_dynamics.py:274,497 and batch_tsa.py:225); ``steer_fused`` runs all rows and all steps in one launch.
"""
import time

import numpy as np
import torch

from ... import ops
from .tsa import _rollout

MAX_FUSED_BATCH = 128     # cbfrrt_rollout_segments: the rows of a batch share one CTA


def _gain_host(dm):
    key = (id(dm.K), dm.K._version)
    if getattr(dm, "_K_np_key", None) != key:
        dm._K_np = dm.K.detach().cpu().numpy().astype(np.float32)
        dm._K_np_key = key
    return dm._K_np


_ZERO_NETS = {}


def _zero_net(n):
    if n not in _ZERO_NETS:
        _ZERO_NETS[n] = np.zeros(ops.packed_weight_count(n, 48, 2), np.float32)
    return _ZERO_NETS[n]


def _dev_index(device):
    d = torch.device(device)
    if d.type != "cuda":
        return 0
    return d.index if d.index is not None else torch.cuda.current_device()


def explore_nearest(non_terminal_states, sample_states, batch, device="cuda"):
    """The nearest-neighbour selection of global_explore (batch_tsa.py:129-145) on the GPU (cbfrrt::knn):
      * tree smaller than ``batch`` (or freshly sampled states): one nearest node per sample (np.argmin per row);
      * otherwise: the ``batch`` nearest nodes of sample_states[0] (np.argsort(dists)[:batch]), the sample repeated.
    -> (nearest_idxs int64 (B,), min_dists float32 (B,), sample_states (B, n)) with the reference's shapes.
    Ties go to the lowest node index, as np.argmin / a stable sort do."""
    tree = np.asarray(non_terminal_states, dtype=np.float32)
    samples = np.asarray(sample_states, dtype=np.float32).reshape(-1, tree.shape[1])
    if tree.shape[0] < batch:
        idx, dist = ops.knn_host(samples, tree, 1, device=_dev_index(device))
        return idx[:, 0].astype(np.int64), dist[:, 0], np.asarray(sample_states)
    idx, dist = ops.knn_host(samples[:1], tree, int(batch), device=_dev_index(device))
    rep = np.array([np.asarray(sample_states)[0] for _ in range(batch)])
    return idx[0].astype(np.int64), dist[0], rep


def explore_nearest_each(non_terminal_states, sample_states, device="cuda"):
    """first branch of global_explore (batch_tsa.py:129-137): one nearest node for every freshly sampled state"""
    tree = np.asarray(non_terminal_states, dtype=np.float32)
    samples = np.asarray(sample_states, dtype=np.float32).reshape(-1, tree.shape[1])
    idx, dist = ops.knn_host(samples, tree, 1, device=_dev_index(device))
    return idx[:, 0].astype(np.int64), dist[:, 0]


@torch.no_grad()
def steer(u, dynamics_model, sample_states, nearests, RRT_step=10, device="cpu"):
    """batch_tsa.py:191-231 -> (x_last (B,n), no_collisions [bool]*B, time_dict, xs_list)."""
    time_dict = {"complete_sample": 0., "cl_dynamics": 0.}
    t0 = time.time()
    B = len(sample_states)
    no_collisions = [True for _ in range(B)]
    xs_list = [[] for _ in range(B)]
    controller_update_freq = dynamics_model.controller_dt // dynamics_model.dt
    dynamics_model.set_intermediate_goals(np.asarray(sample_states))
    x_current = dynamics_model.complete_sample_with_observations(
        torch.tensor(np.asarray(nearests), device=device, dtype=torch.float32), len(nearests))
    time_dict["complete_sample"] += (time.time() - t0) / B
    for tstep in range(RRT_step):
        if tstep % controller_update_freq == 0:
            u_current = u(x_current)
            if isinstance(u_current, tuple):
                u_current, u_time_dict = u_current
            else:
                u_time_dict = {}
            for key in u_time_dict.keys():
                time_dict[key] = time_dict.get(key, 0.) + u_time_dict[key] / B
        t1 = time.time()
        x_current = dynamics_model.closed_loop_dynamics(
            x_current, u_current, update_observation=((tstep + 1) % controller_update_freq == 0))
        time_dict["cl_dynamics"] += (time.time() - t1) / B
        unsafe = dynamics_model.unsafe_mask(x_current).cpu().numpy()  # one launch instead of B (batch_tsa.py:225)
        x_np = x_current.cpu().numpy()
        for i in range(B):
            stuck = len(xs_list[i]) > 10 and np.linalg.norm(xs_list[i][-1] - xs_list[i][-10]) < 3e-2
            no_collisions[i] = no_collisions[i] and not stuck and not bool(unsafe[i])
            if no_collisions[i]:
                xs_list[i].append(x_np[i])
        if True not in no_collisions:
            break
    return x_current.cpu().numpy()[:, :dynamics_model.q_dims], no_collisions, time_dict, xs_list


@torch.no_grad()
def steer_fused(controller, dynamics_model, sample_states, nearests, RRT_step=10, want_traj=False):
    """All rows x all steps in one launch (no stuck test, like the timed closed-loop rollouts of BASELINE
    config 4) -> (q_final (B,n) tensor, alive bool (B,), steps_done (B,), traj or None)."""
    dm = dynamics_model
    dm.set_intermediate_goals(np.asarray(sample_states))
    q0 = torch.as_tensor(np.asarray(nearests), dtype=torch.float32).reshape(-1, dm.q_dims).to(dm.device)
    qf, alive, done, traj = _rollout(controller, dm, q0, RRT_step, want_traj=want_traj)
    return qf, alive.bool(), done, (traj if want_traj else None)


@torch.no_grad()
def steer_fused_reference(controller, dynamics_model, sample_states, nearests, RRT_step=10):
    """The batched steer of the reference (batch_tsa.py:191-231) with ITS semantics in one launch
    -> (x_last (B,n) numpy, no_collisions [bool]*B, xs_list) exactly as ``steer(controller.u, ...)`` returns them:

      * the stuck test runs before the append, on the stored list (more than 10 states: xs[-1] vs xs[-10], 3e-2), and a
        stuck row is reported as a collision (``no_collisions[i] = False``, batch_tsa.py:226-227);
      * a dead row (stuck or unsafe) never appends again, but its state keeps integrating: x_last is the running
        x_current of every row, dead or alive (batch_tsa.py:231);
      * the loop ends early once EVERY row is dead (batch_tsa.py:229-230): x_last is then the state after that step.
        That is the one batch-wide dependency of the reference's loop; it is reproduced by a second launch truncated at
        the step where the last row died (rare: all rows of a batch must die).
    Kernel: cbfrrt_rollout with stuck_mode = 1, keep_going = 1 (include/cbfrrt.h)."""
    dm = dynamics_model
    dm.set_intermediate_goals(np.asarray(sample_states))
    q0 = torch.as_tensor(np.asarray(nearests), dtype=torch.float32).reshape(-1, dm.q_dims).to(dm.device)
    steps = int(RRT_step)
    qf, alive, done, traj = _rollout(controller, dm, q0, steps, stuck=(9, 10, 3e-2), want_traj=True, keep_going=True,
                                     stuck_mode=1)
    alive_np, done_np = alive.cpu().numpy().astype(bool), done.cpu().numpy()
    if steps > 0 and not alive_np.any():
        last_death = int(done_np.max())          # a row that died at step t (0-based) has appended exactly t states
        if last_death + 1 < steps:
            qf, alive, done, traj = _rollout(controller, dm, q0, last_death + 1, stuck=(9, 10, 3e-2), want_traj=True,
                                             keep_going=True, stuck_mode=1)
            alive_np, done_np = alive.cpu().numpy().astype(bool), done.cpu().numpy()
    traj_np = traj.cpu().numpy()
    xs_list = [[row for row in traj_np[i, :done_np[i]]] for i in range(traj_np.shape[0])]
    return qf.cpu().numpy(), [bool(a) for a in alive_np], xs_list


@torch.no_grad()
def steer_segments_fused(controller, dynamics_model, sample_states, nearests, n_segments, RRT_step=20, nominal=False):
    """The segment loop of global_explore (batch_tsa.py:154-170: ``RRT_PARAM // 20`` batched steer calls, each starting
    where the previous one ended) in ONE launch and ONE read-back (cbfrrt_rollout_segments), with the reference's
    semantics per segment (batch_tsa.py:191-231, incl. the batch-wide early exit once every row is dead).
    -> list over segments of (x_last (B,n) numpy, no_collisions [bool]*B, xs_list), exactly what ``n_segments``
    consecutive ``steer(controller.u, ...)`` calls return.  B <= MAX_FUSED_BATCH.
    ``nominal``: the 'line' steer of the vanilla planners, ``steer(dynamics_model.u_nominal, ...)`` (batch_tsa.py:158): the
    same kernel with relaxation penalty 0 -- the QP's multiplier is confined to [0, penalty], so u is exactly the clamped
    LQR control u_nominal (arm_dynamics.py:124-161) and the network is never consulted (a zero network is passed)."""
    dm = dynamics_model
    goals = np.asarray(sample_states, dtype=np.float32).reshape(-1, dm.q_dims)
    dm.set_intermediate_goals(np.asarray(sample_states))
    q0 = np.asarray(nearests, dtype=np.float32).reshape(-1, dm.q_dims)
    if goals.shape[0] == 1 and q0.shape[0] > 1:
        goals = np.repeat(goals, q0.shape[0], 0)
    env = dm.env
    if nominal:
        upper, lower = dm.control_limits
        lo, hi = lower.detach().cpu().numpy().astype(np.float32), upper.detach().cpu().numpy().astype(np.float32)
        weights, hidden, layers, lam, penalty = _zero_net(dm.q_dims), 48, 2, 0.0, 0.0
    else:
        lo, hi = controller._limits_host()
        weights, hidden, layers = controller._weights_host(), controller.h_hidden_size, controller.h_hidden_layers
        lam, penalty = float(controller.clf_lambda), float(controller.clf_relaxation_penalty)
    # numpy in / numpy out (cbfrrt_rollout_segments_host): one pinned H2D copy, the launch, one D2H copy
    seg_q, seg_alive, seg_done, traj = ops.rollout_segments_host(
        dm.robot.name, q0, goals, int(n_segments), int(RRT_step), int(dm.controller_dt // dm.dt), float(dm.dt), 9, 10, 3e-2,
        np.concatenate([env.obs_positions, env.obs_sizes], 1), env.obs_spheres, bool(env.include_floor), weights, hidden, layers,
        _gain_host(dm), lo, hi, lam, penalty, float(dm.dis_threshold), float(dm.thr_hard), device=_dev_index(dm.device))
    out = []
    for k in range(int(n_segments)):
        xs = [[row for row in traj[i, k * RRT_step:k * RRT_step + seg_done[i, k]]] for i in range(traj.shape[0])]
        out.append((seg_q[:, k], [bool(a) for a in seg_alive[:, k]], xs))
    return out
