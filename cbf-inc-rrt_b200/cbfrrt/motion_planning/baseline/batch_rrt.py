"""Batched RRT front-end -- mirror of motion_planning/baseline/batch_tsa.py:11-189 and search_tree.py:6-92.

The planner's control flow (goal bias, uniform sampling, nearest-node selection, steer, insertion) is the
reference's; what changed is where the work runs:
  * nearest / k-nearest selection (batch_tsa.py:129-145)     -> cbfrrt::knn         (explore_nearest)
  * the steer loops of a whole batch (batch_tsa.py:157-166)   -> cbfrrt::rollout     (one launch per 20 sim steps)
  * goal test of the new states (batch_tsa.py:171)            -> one batched safe_mask (cbfrrt::masks)
Tree bookkeeping stays host Python like the reference's, but appends are amortised (the reference re-allocates the
state arrays with np.append on every insertion, search_tree.py:70-71).
"""
import time

import numpy as np
import torch

from .batch_tsa import (explore_nearest, explore_nearest_each, steer as batch_steer, steer_fused_reference,
                        steer_segments_fused, MAX_FUSED_BATCH)
from .tsa import _rollout


class _Grow:
    """append-only row store: numpy view on a doubling buffer"""

    def __init__(self, first):
        first = np.asarray(first, dtype=np.float64).reshape(1, -1)
        self.buf, self.n = np.empty((64, first.shape[1])), 1
        self.buf[0] = first

    def append(self, row):
        if self.n == self.buf.shape[0]:
            self.buf = np.concatenate([self.buf, np.empty_like(self.buf)])
        self.buf[self.n] = row
        self.n += 1

    @property
    def view(self):
        return self.buf[:self.n]


class SearchTree:
    """search_tree.py:6-48: same attributes (states, nominal_states, parents, rewired_parents, freesp, x_list,
    in_goal_region, non_terminal_states, non_terminal_idxes, ...) and the same path() result."""

    def __init__(self, root):
        self._states, self._nominal, self._nonterm = _Grow(root), _Grow(root), _Grow(root)
        self.parents, self.rewired_parents = [None], [None]
        self.expanded_by_rrt, self.freesp = [None], [True]
        self.x_list = [[np.asarray(root)]]
        self.costs, self.path_lengths = [0.], [-1]
        self.cumulated_collision_checks = [0]
        self.in_goal_region = [False]
        self.non_terminal_idxes = [0]

    states = property(lambda self: self._states.view)
    nominal_states = property(lambda self: self._nominal.view)
    non_terminal_states = property(lambda self: self._nonterm.view)

    def path(self, idx=None):
        """(states, accumulated negative edge lengths, nominal states, edge state lists) from the root to node ``idx``
        (default: the LAST inserted node, search_tree.py:24-47) if that node is in the goal region, else four empty
        lists."""
        idx = len(self.parents) - 1 if idx is None else idx
        if not self.in_goal_region[idx]:
            return [], [], [], []
        path, nominal, costs, xs = [], [], [], []
        cost = 0
        while True:
            path.append(self.states[idx])
            nominal.append(self.nominal_states[idx])
            xs.append(self.x_list[idx])
            costs.append(cost)
            if idx == 0:
                break
            parent = self.rewired_parents[idx]
            cost -= np.linalg.norm(self.states[idx] - self.states[parent])
            idx = parent
        return path[::-1], costs[::-1], nominal[::-1], xs[::-1]


def insert_new_state(search_tree, state, nominal_state, x_list, parent_idx, no_collision, done, expanded_by_rrt=False):
    """search_tree.py:68-92 -> index of the new node"""
    t = search_tree
    t._states.append(state)
    t._nominal.append(nominal_state)
    t.parents.append(parent_idx)
    t.rewired_parents.append(parent_idx)
    t.x_list.append(x_list)
    t.expanded_by_rrt.append(expanded_by_rrt)
    t.freesp.append(no_collision)
    t.in_goal_region.append(done)
    t.path_lengths.append(t.path_lengths[-1])
    t.costs.append(-1)
    idx = t._states.n - 1
    if not done:
        t._nonterm.append(state)
        t.non_terminal_idxes.append(idx)
    return idx


def _steer_batch(dynamics_model, controller, steer_type, sample_states, current_states, rrt_step):
    """-> (new_states (B,n) numpy, no_collisions list[bool], edge_states list of lists)"""
    dm = dynamics_model
    if steer_type == "line":      # nominal controller: the per-step host loop on batched ops (batch_tsa.py:158)
        new_states, ok, _, edges = batch_steer(dm.u_nominal, dm, sample_states, current_states, RRT_step=rrt_step,
                                               device=dm.device)
        return new_states, ok, edges
    if steer_type != "cbf":
        raise ValueError("Unknown steer_type, must be one of {'line', 'cbf'}.")
    assert hasattr(controller, "u")
    # the reference's batched steer semantics (batch_tsa.py:191-231: stuck rows count as collisions, dead rows keep
    # integrating, x_last = running state) in one launch -- same nodes and flags as the 'line' branch's host loop
    return steer_fused_reference(controller, dm, sample_states, current_states, RRT_step=rrt_step)


def _goal_test(dynamics_model, states):
    """dynamics_model.goal_mask (arm_dynamics.py:234-262: within 0.2 rad of the goal AND safe) for the new states of a
    batch: the distance part is evaluated first, in the same float32 torch expression, and the safe_mask launch is made
    only for the rows inside the goal ball -- none, for most expansions."""
    dm = dynamics_model
    xs = np.asarray(states, dtype=np.float32).reshape(-1, dm.q_dims)
    # screen in numpy (an expansion makes this test for a handful of rows): a row whose distance is not within 1e-4 of the
    # 0.2 rad threshold gets the same answer from any float32 evaluation order; otherwise the reference's own torch
    # expression decides
    d = np.sqrt(((xs - np.asarray(dm.goal_state[:dm.q_dims], dtype=np.float32)) ** 2).sum(1))
    if not (d <= 0.2 + 1e-4).any():
        return np.zeros(xs.shape[0], dtype=bool)
    x = torch.from_numpy(xs)
    goal = torch.as_tensor(dm.goal_state[:dm.q_dims]).type_as(x)
    near = (x[:, :dm.q_dims] - goal).norm(dim=1) <= 0.2
    done = torch.zeros(x.shape[0], dtype=torch.bool)
    if bool(near.any()):
        done[near] = dm.safe_mask(x[near]).cpu()
    return done.numpy()


def global_explore(search_tree, dynamics_model, sample_states=None, steer_type="line", controller=None, batch=50,
                   RRT_PARAM=20, rng=None):
    """One batched RRT expansion (batch_tsa.py:99-189).  Yields, per expanded sample,
    (new_state, sample_state, parent node index, no_collision, done, time_dict).  ``rng``: np.random or a
    np.random.RandomState (rand / randint), for reproducible runs."""
    rng = rng if rng is not None else np.random
    t0 = time.time()
    dm = dynamics_model
    tree_states = search_tree.non_terminal_states
    if sample_states is None:
        RRT_PARAM = 120
        sample_states = dm.sample_state_space(batch).cpu().numpy()[:, :dm.q_dims]
        nearest_idxs, _ = explore_nearest_each(tree_states, sample_states, device=dm.device)
    else:
        nearest_idxs, _, sample_states = explore_nearest(tree_states, sample_states, batch, device=dm.device)
    if steer_type == "cbf":       # occasionally expand a random node instead of the nearest (batch_tsa.py:147-152)
        for i in range(min(len(nearest_idxs), batch, tree_states.shape[0])):
            if rng.rand() > 0.9:
                nearest_idxs[i] = rng.randint(0, tree_states.shape[0])
    t1 = time.time()
    assert RRT_PARAM % 20 == 0
    current = tree_states[nearest_idxs]
    first_parents = [search_tree.non_terminal_idxes[j] for j in nearest_idxs]
    parents = list(first_parents)
    new_states, no_collisions = current, [True] * len(parents)
    n_seg = RRT_PARAM // 20
    if steer_type not in ("line", "cbf"):
        raise ValueError("Unknown steer_type, must be one of {'line', 'cbf'}.")
    fused = n_seg >= 1 and len(parents) <= MAX_FUSED_BATCH
    if fused:   # all segments of the expansion in one launch and one read-back (cbfrrt_rollout_segments)
        assert steer_type == "line" or hasattr(controller, "u")
        segments = steer_segments_fused(controller, dm, sample_states, current, n_seg, RRT_step=20, nominal=steer_type == "line")
    if fused:   # one goal test for the new states of all segments (same float32 torch expression, one call instead of n_seg)
        done_all = _goal_test(dm, np.concatenate([seg[0] for seg in segments])).reshape(n_seg, -1)
    for k in range(n_seg):
        if fused:
            new_states, no_collisions, edges = segments[k]
            done = done_all[k]
        else:
            new_states, no_collisions, edges = _steer_batch(dm, controller, steer_type, sample_states, current, 20)
            done = _goal_test(dm, new_states)
        for i in range(len(parents)):
            parents[i] = insert_new_state(search_tree, new_states[i], sample_states[i], edges[i], parents[i],
                                          no_collisions[i], bool(done[i]), expanded_by_rrt=True)
        current = new_states
    t2 = time.time()
    if RRT_PARAM // 20 == 0:
        done = _goal_test(dm, new_states)    # (otherwise: the goal test of the last segment, same states)
    for i in range(len(parents)):
        t_dict = {"explore_1": (t1 - t0) / len(parents), "total_steer": (t2 - t1) / len(parents),
                  "total_explore": (time.time() - t0) / len(parents)}
        yield new_states[i], sample_states[i], first_parents[i], no_collisions[i], bool(done[i]), t_dict


def batch_RRT_plan(env, dynamics_model, init_state, goal_state, RRT_PARAM, controller=None, T=100, g_explore_eps=1.,
                   optimal_version=False, stop_when_success=False, model_eps=0.05, steer_type="line", batch=50, rng=None):
    """batch_tsa.py:11-97: same arguments, same result dict (success, path, edge_states, nominal_path, path_cost,
    explored_nodes, time_dict)."""
    rng = rng if rng is not None else np.random
    if optimal_version:
        raise ValueError("Optimal version motion planning not supported.")
    start = time.time()
    tree = SearchTree(root=np.asarray(init_state, dtype=np.float64))
    time_dict, success, i = {}, False, 0
    while i < T:
        if rng.rand() < model_eps:            # goal bias
            samples = np.array([goal_state])
        elif rng.rand() < g_explore_eps:      # uniform exploration
            samples = None
        else:
            raise ValueError("g_explore_eps set less than 1, guided expansion not supported at present.")
        for _, _, _, _, done, td in global_explore(tree, dynamics_model, sample_states=samples, steer_type=steer_type,
                                                   controller=controller, batch=batch, RRT_PARAM=RRT_PARAM, rng=rng):
            i += 1
            success = success or done
            for k, v in td.items():
                time_dict[k] = time_dict.get(k, 0.) + v
            if success and stop_when_success:
                break
        if success and stop_when_success:
            break
    time_dict["total_time"] = time.time() - start
    # the reference reports the path to the LAST inserted node (search_tree.py:25); a batch inserts its other members
    # after the one that reached the goal, so fall back to the first goal node
    path = tree.path()
    if success and not path[0]:
        path = tree.path(tree.in_goal_region.index(True))
    return {"success": success, "path": path[0], "edge_states": path[3], "nominal_path": path[2], "path_cost": path[1],
            "explored_nodes": i, "time_dict": time_dict, "search_tree": tree}
