"""CPU suite: host logic of the reference-shaped API (cbfrrt.environment / neural_cbf / motion_planning) with the
CUDA ops replaced by oracle-backed fakes (tests/fake_ops.py).  Checks the drop-in contract of SURVEY.md 8b:
signatures, shapes, dtypes, devices, return conventions and the control flow of steer / edge_checking."""
import numpy as np
import pytest
import torch

import fake_ops


@pytest.fixture()
def stack(monkeypatch, coracle):
    fake_ops.install(monkeypatch)
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController

    def make(robot_name="panda"):
        env = ArmEnv([robot_name], config_file="", include_floor=True, device="cpu")
        robot = env.robot_list[0]
        dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=robot)
        torch.manual_seed(1)
        ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0,
                                         safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48,
                                         use_neural_actor=False)
        return env, robot, dyn, ctrl
    return make


def test_env_and_robot_attributes(stack):
    env, robot, dyn, ctrl = stack("panda")
    assert env.obstacle_ids == [0, 1, 2, 3] and env.obs_positions.shape == (3, 3) and env.include_floor
    assert np.allclose(env.get_env_config(-1)[0][0], [0.3, 0.17, 0.3])               # arm_env.py:113-114
    assert robot.body_joints == list(range(7)) and robot.ee_joints == [9, 10] and robot.n_joints == 12
    assert robot.body_dim == 7 and robot.ee_dim == 2 and robot.body_range.shape == (7, 2)
    assert str(robot) == "Franka_Panda" and robot.invalid_link == [7, 11]
    env.reset_env((np.array([[0.4, 0.0, 0.5]]), np.array([[0.1, 0.1, 0.1]])))
    assert env.obstacle_ids == [0, 1] and env.scene[0].shape == (1, 8)
    with pytest.raises(NotImplementedError):
        ArmEnv = type(env)
        ArmEnv(["iiwa"], device="cpu")                                                 # arm_env.py:79-80
    env2, robot2, dyn2, _ = stack("magician")
    assert robot2.body_dim == 4 and str(robot2) == "Magician" and np.allclose(robot2.q0, [0, 0.707, 0.707, 0])
    with pytest.raises(AttributeError):
        env.p.getClosestPoints()


def test_dynamics_properties_and_masks(stack, coracle):
    env, robot, dyn, ctrl = stack("panda")
    assert (dyn.n_dims, dyn.q_dims, dyn.n_controls, dyn.o_dims, dyn.state_aux_dims) == (7, 7, 7, 1, 7)
    upper, lower = dyn.state_limits                                                    # (upper, lower) order
    assert (upper > lower).all() and abs(upper[5].item() - 3.8223) < 1e-6
    cu, cl = dyn.control_limits
    assert torch.equal(cl, -cu) and abs(cu[0].item() - 2.175) < 1e-6
    assert abs(dyn.K[0, 0].item() - 0.9834722117237553) < 1e-12  # reference value: B is float32 (golden K) and dyn.K.shape == (7, 7)
    assert str(dyn) == "Franka_Panda_mindis_Dynamics"
    x = dyn.sample_state_space(64)
    assert x.shape == (64, 7) and (x <= upper).all() and (x >= lower).all()
    safe, unsafe, bnd = dyn.safe_mask(x), dyn.unsafe_mask(x), dyn.boundary_mask(x)
    assert safe.dtype == torch.bool and safe.shape == (64,) and torch.equal(bnd, ~(safe | unsafe))
    assert not (safe & unsafe).any()
    datax = dyn.complete_sample_with_observations(x, 64)
    assert datax.shape == (64, 15) and torch.equal(datax[:, :7], x)
    o = dyn._get_observation_with_state(x[0])
    assert o.shape == (8,) and abs(o[0] - datax[0, 7].item()) < 1e-7
    # goal mask: within 0.2 of the goal AND safe (arm_dynamics.py:234-262)
    g = torch.tensor(robot.q0, dtype=torch.float32).reshape(1, -1)
    assert dyn.goal_mask(g).item() == dyn.safe_mask(g).item()
    assert not dyn.goal_mask(g + 0.2).item()
    xs = dyn.sample_safe(5)
    assert xs.shape == (5, 15) and dyn.safe_mask(xs).all()
    with pytest.raises(RuntimeWarning):
        dyn.sample_with_mask(5, lambda s: torch.zeros(s.shape[0], dtype=torch.bool), max_tries=2)


def test_closed_loop_dynamics_conventions(stack):
    env, robot, dyn, ctrl = stack("panda")
    x = dyn.complete_sample_with_observations(dyn.sample_state_space(4), 4)
    u = torch.ones(4, 7) * 0.5
    nxt, times = dyn.closed_loop_dynamics(x, u, update_observation=False, return_time=True)
    assert len(times) == 2 and nxt.shape == (4, 15)
    assert torch.allclose(nxt[:, :7], x[:, :7] + 0.5 / 120)                           # returns x_next, not xdot
    assert (nxt[:, 7:] == 0).all()                                                     # stale o/aux are zeros
    nxt2 = dyn.closed_loop_dynamics(x, u, update_observation=True)
    assert torch.equal(nxt2[:, :7], nxt[:, :7]) and (nxt2[:, 7] != 0).any()
    f, g = dyn.control_affine_dynamics(x[:, :7])
    assert f.shape == (4, 7, 1) and not f.any() and torch.equal(g[0], torch.eye(7))
    dyn.set_goal(np.zeros(7))
    dyn.set_intermediate_goals(np.ones((4, 7)))
    un = dyn.u_nominal(x)
    cu, cl = dyn.control_limits
    assert torch.allclose(un, torch.clamp(-(x[:, :7] - 1.0) * dyn.K[0, 0].float(), cl, cu), atol=1e-6)


def test_controller_conventions_and_reference_autograd(stack, golden):
    env, robot, dyn, ctrl = stack("panda")
    from conftest import golden_state_dict, rel_err
    g = golden["panda"]
    ctrl.load_weights({"h_nn." + k: torch.tensor(v) for k, v in golden_state_dict(g).items()})
    assert set(ctrl.state_dict()) == {"h_nn." + k for k in golden_state_dict(g)}       # checkpoint key names
    dyn.set_goal(g["shared_goal"])
    datax = torch.tensor(g["datax"])
    u, t_dict = ctrl.u(datax)                                                          # (u, t_dict) tuple
    assert isinstance(t_dict, dict) and {"V_w_Jacobian", "lie_derivative", "nn_forward", "QP"} <= set(t_dict)
    assert rel_err(u.numpy(), g["shared_u"]) < 1e-5
    h, J, extra = ctrl.h_with_jacobian(datax)
    assert h.shape == (96, 1) and J.shape == (96, 1, 7) and extra == {}
    assert rel_err(h.numpy(), g["shared_h"], floor=0.1) < 1e-5 and rel_err(J.numpy(), g["shared_J"], floor=0.1) < 1e-5
    assert torch.equal(ctrl.h(datax[:, :8]), h) and torch.equal(ctrl.V(datax), h)
    V, Lf, Lg, td = ctrl.V_with_lie_derivatives(datax)
    assert Lf.shape == (96, 1, 1) and not Lf.any() and torch.equal(Lg, J)
    (u2, r2), td = ctrl.solve_CLF_QP(datax, relaxation_penalty=0.05)
    assert r2.shape == (96, 1) and rel_err(u2.numpy(), g["shared_u_pen0.05"]) < 1e-5
    uref = torch.zeros(96, 7)
    (u3, _), _ = ctrl.solve_CLF_QP(datax, u_ref=uref)
    inactive = (0.1 * h[:, 0] <= 0)
    assert torch.equal(u3[inactive], uref[inactive])
    with pytest.raises(AssertionError):
        ctrl.solve_CLF_QP(datax, u_ref=torch.zeros(3, 7))
    # per-row goals (batch_tsa.py:199)
    dyn.set_intermediate_goals(g["rows_goal"])
    u4, _ = ctrl.u(datax)
    assert rel_err(u4.numpy(), g["rows_u"]) < 1e-5


def test_generic_clf_controller_qp(stack):
    """CLFController / CBFController with the quadratic V (clf_controller.py:141-167, cbf_controller.py:53-67)."""
    env, robot, dyn, _ = stack("magician")
    from cbfrrt.neural_cbf.controllers import CBFController
    c = CBFController(dyn, [{"m1": 1.0}], None, cbf_lambda=1.0, cbf_relaxation_penalty=50.0, controller_period=1 / 30)
    x = dyn.sample_state_space(32)
    dyn.set_goal(np.zeros(4))
    V, JV, _ = c.V_with_jacobian(x)
    P = dyn.P.float()
    assert torch.allclose(V, 0.5 * ((x @ P) * x).sum(1) - 1.0, atol=1e-5) and JV.shape == (32, 1, 4)
    u, td = c.u(x)
    Vv, Lf, Lg, _ = c.V_with_lie_derivatives(x)
    cu, cl = dyn.control_limits
    assert (u <= cu + 1e-6).all() and (u >= cl - 1e-6).all()
    (uu, r), _ = c.solve_CLF_QP(x)
    g = Lf[:, 0, 0] + (Lg[:, 0] * uu).sum(1) + 1.0 * Vv - r[:, 0]
    assert (g < 1e-4).all() and (r >= 0).all()


def test_clf_controller_multi_scenario(stack):
    """len(scenarios) > 1: one relaxed constraint per scenario (clf_controller.py:86-139); r has one column each and the
    control satisfies every scenario's constraint up to its own relaxation."""
    env, robot, dyn, _ = stack("magician")
    from cbfrrt.neural_cbf.controllers import CBFController
    scen = [{"m1": 1.0}, {"m1": 1.2}, {"m1": 0.8}]
    c = CBFController(dyn, scen, None, cbf_lambda=1.0, cbf_relaxation_penalty=50.0, controller_period=1 / 30)
    x = dyn.sample_state_space(48)
    dyn.set_goal(np.zeros(4))
    (u, r), _ = c.solve_CLF_QP(x)
    assert u.shape == (48, 4) and r.shape == (48, 3)
    V, Lf, Lg, _ = c.V_with_lie_derivatives(x)
    g = Lf[:, :, 0] + torch.einsum("bsk,bk->bs", Lg, u) + V.reshape(-1, 1) - r
    assert (g < 1e-4).all() and (r >= 0).all()
    # identical scenarios (ArmDynamics ignores params): the answer equals the single-scenario controller's
    c1 = CBFController(dyn, scen[:1], None, cbf_lambda=1.0, cbf_relaxation_penalty=50.0 * 3, controller_period=1 / 30)
    (u1, r1), _ = c1.solve_CLF_QP(x)
    assert torch.allclose(u, u1, atol=2e-5)


def test_steer_reference_loop_equals_fused(stack):
    env, robot, dyn, ctrl = stack("panda")
    from cbfrrt.motion_planning.baseline import steer, steer_fused, batch_steer, batch_steer_fused
    x = dyn.sample_state_space(200)
    safe = dyn.safe_mask(x)
    xs = x[safe]
    n_coll = 0
    for i in range(6):
        start, target = xs[2 * i].numpy(), xs[2 * i + 1].numpy()
        x_last, ok, td, x_list = steer(ctrl.u, dyn, target, start, RRT_step=30, device="cpu")
        x_last2, ok2, td2, x_list2 = steer_fused(ctrl, dyn, target, start, RRT_step=30)
        assert ok == ok2 and len(x_list) == len(x_list2)
        assert np.abs(np.asarray(x_list[1:]) - np.asarray(x_list2[1:])).max() < 1e-5 if len(x_list) > 1 else True
        assert {"complete_sample", "cl_dynamics", "collision_checking", "nn_forward", "QP"} <= set(td)
        assert np.array_equal(x_list[0], start) and 2 <= len(x_list) <= 31
        n_coll += (not ok)
    # nominal (line) steering uses the same loop with dynamics_model.u_nominal (tsa.py:179-181)
    x_last, ok, td, x_list = steer(dyn.u_nominal, dyn, xs[1].numpy(), xs[0].numpy(), RRT_step=30, device="cpu")
    assert len(x_list) >= 2 and "u" in td
    # batched steer: same rows as the fused launch while nothing collides / sticks
    starts, targets = xs[:8].numpy(), xs[8:16].numpy()
    xb, oks, tdb, xs_list = batch_steer(ctrl.u, dyn, targets, starts, RRT_step=8, device="cpu")
    qf, alive, done, _ = batch_steer_fused(ctrl, dyn, targets, starts, RRT_step=8)
    assert xb.shape == (8, 7) and len(oks) == 8 and len(xs_list) == 8
    keep = np.array(oks) & alive.numpy()
    assert keep.sum() >= 4 and np.abs(xb[keep] - qf.numpy()[keep]).max() < 1e-5


def test_batched_steer_reference_semantics_fused(stack):
    """the one-launch batched steer with the reference's semantics (batch_tsa.py:191-231: stuck rows become
    collisions, dead rows keep integrating, early exit when the whole batch is dead) against the host-loop mirror that
    the reference's own batch_tsa.steer is pinned to (test_reference_planner_cpu)"""
    env, robot, dyn, ctrl = stack("panda")
    from cbfrrt.motion_planning.baseline import batch_steer, batch_steer_fused_reference
    x = dyn.sample_state_space(400)
    xs = x[dyn.safe_mask(x)].numpy()[:, :7]
    assert xs.shape[0] >= 40
    seen = {"dead": 0, "stuck": 0, "all_dead": 0}
    cases = [(xs[:12], xs[12:24], 40), (xs[24:36], xs[24:36] + 0.02, 40),            # ordinary / targets next to the start (stuck)
             (xs[:6], np.repeat(np.array([[0., 1.7, 0., -0.1, 0., 3.7, 0.]], np.float32), 6, 0), 60)]  # a pose that folds onto the floor
    for starts, targets, steps in cases:
        xb, oks, _, xs_list = batch_steer(ctrl.u, dyn, targets, starts, RRT_step=steps, device="cpu")
        xf, okf, lists_f = batch_steer_fused_reference(ctrl, dyn, targets, starts, RRT_step=steps)
        assert list(oks) == list(okf)
        assert [len(a) for a in xs_list] == [len(b) for b in lists_f]
        assert np.abs(xb - xf).max() < 1e-4
        for a, b in zip(xs_list, lists_f):
            if len(a):
                assert np.abs(np.asarray(a)[:, :7] - np.asarray(b)).max() < 1e-4
        seen["dead"] += sum(not o for o in oks)
        seen["all_dead"] += int(not any(oks))
        # a stuck row: died although its next state was not unsafe -- identify via the stored list (>= 11 states, last two 9 apart close)
        for ok, lst in zip(oks, xs_list):
            if not ok and len(lst) > 10 and np.linalg.norm(np.asarray(lst[-1]) - np.asarray(lst[-10])) < 3e-2:
                seen["stuck"] += 1
    assert seen["dead"] > 0 and seen["stuck"] > 0, seen


def test_fused_segments_equal_consecutive_batched_steers(stack):
    """cbfrrt_rollout_segments (one launch for all segments of an expansion) against what global_explore did before:
    n_seg consecutive calls of the batched steer host loop, each starting where the previous one ended
    (batch_tsa.py:154-170), incl. batches of ONE row (the batch-wide early exit fires at every death) and rows that die"""
    env, robot, dyn, ctrl = stack("panda")
    from cbfrrt.motion_planning.baseline import batch_steer, steer_segments_fused
    x = dyn.sample_state_space(400)
    xs = x[dyn.safe_mask(x)].numpy()[:, :7]
    fold = np.array([[0., 1.7, 0., -0.1, 0., 3.7, 0.]], np.float32)            # a pose that folds onto the floor
    n_dead = n_early = 0
    for starts, targets in ((xs[:5], xs[5:10]), (xs[10:11], fold), (xs[11:14], np.repeat(fold, 3, 0)), (xs[20:21], xs[20:21] + 0.02),
                            (xs[30:36], xs[36:42])):
        n_seg = 4
        segs = steer_segments_fused(ctrl, dyn, targets, starts, n_seg, RRT_step=20)
        cur = starts
        for k in range(n_seg):
            xb, oks, _, lists = batch_steer(ctrl.u, dyn, targets, cur, RRT_step=20, device="cpu")
            xf, okf, lf = segs[k]
            assert list(oks) == list(okf), (k, oks, okf)
            assert [len(a) for a in lists] == [len(b) for b in lf]
            assert np.abs(xb - xf).max() < 2e-4
            for a, b in zip(lists, lf):
                if len(a):
                    assert np.abs(np.asarray(a)[:, :7] - np.asarray(b)).max() < 2e-4
            n_dead += sum(not o for o in oks)
            n_early += int(not any(oks) and max(len(a) for a in lists) < 19)
            cur = xb
    assert n_dead > 0 and n_early > 0, (n_dead, n_early)


def test_fused_line_steer_equals_nominal_host_loop(stack):
    """the vanilla planners' 'line' steer (steer(dynamics_model.u_nominal, ...), batch_tsa.py:158) through the same fused
    kernel with relaxation penalty 0: nodes, flags and stored states of the per-step host loop"""
    env, robot, dyn, ctrl = stack("panda")
    from cbfrrt.motion_planning.baseline import batch_steer, steer_segments_fused
    x = dyn.sample_state_space(400)
    xs = x[dyn.safe_mask(x)].numpy()[:, :7]
    fold = np.array([[0., 1.7, 0., -0.1, 0., 3.7, 0.]], np.float32)
    n_dead = 0
    for starts, targets in ((xs[:5], xs[5:10]), (xs[10:11], fold), (xs[20:21], xs[20:21] + 0.02), (xs[30:36], xs[36:42])):
        n_seg = 3
        segs = steer_segments_fused(None, dyn, targets, starts, n_seg, RRT_step=20, nominal=True)
        cur = starts
        for k in range(n_seg):
            xb, oks, _, lists = batch_steer(dyn.u_nominal, dyn, targets, cur, RRT_step=20, device="cpu")
            xf, okf, lf = segs[k]
            assert list(oks) == list(okf), (k, oks, okf)
            assert [len(a) for a in lists] == [len(b) for b in lf]
            assert np.abs(xb - xf).max() < 1e-5
            for a, b in zip(lists, lf):
                if len(a):
                    assert np.abs(np.asarray(a)[:, :7] - np.asarray(b)).max() < 1e-5
            n_dead += sum(not o for o in oks)
            cur = xb
    assert n_dead > 0


def test_edge_checking_reference_loop_equals_batched(stack):
    env, robot, dyn, ctrl = stack("panda")
    from cbfrrt.motion_planning.baseline import (edge_checking, edge_checking_batch, edge_checking_eps,
                                                 edge_checking_reference_loop)
    x = dyn.sample_state_space(120)
    xs = x[dyn.safe_mask(x)].numpy()
    eps, agree, n_valid = 0.05, 0, 0
    A = np.stack([xs[i] for i in range(10)])
    Bq = np.stack([xs[i] + 0.5 * (xs[i + 10] - xs[i]) for i in range(10)])
    Bq[3] = A[3] + 1e-3          # K = 0: only the end point is checked (tsa.py:279-281)
    v_eps, fb_eps = edge_checking_eps(A, Bq, dyn, eps)
    for i in range(10):
        a, b = A[i], Bq[i]
        K = int(np.linalg.norm(a - b) / eps)
        ref = edge_checking_reference_loop(a, b, dyn, eps)
        assert edge_checking(a, b, dyn, eps) == bool(v_eps[i])
        if K >= 1:
            v, fb = edge_checking_batch(a[None], b[None], dyn, K)
            assert bool(v.item()) == bool(v_eps[i]) and fb.item() == fb_eps[i].item()
        agree += (bool(v_eps[i]) == ref)
        n_valid += ref
        assert (fb_eps[i].item() == -1) == bool(v_eps[i])
    # the reference interpolates in float64 on the host, the kernel in float32: allow one borderline flip
    assert agree >= 9 and 0 < n_valid < 10


def test_explore_nearest_matches_reference_numpy(stack):
    """explore_nearest == the three numpy branches of global_explore (batch_tsa.py:129-145)"""
    env, robot, dyn, ctrl = stack("panda")
    from cbfrrt.motion_planning.baseline import explore_nearest
    rng = np.random.default_rng(0)
    batch = 50
    samples = dyn.sample_state_space(batch).numpy()
    for n_tree in (7, 400):
        tree = rng.uniform(-1, 1, size=(n_tree, 7)).astype(np.float32)
        idxs, dists, ss = explore_nearest(tree, samples, batch, device="cpu")
        if n_tree < batch:
            d = np.linalg.norm(samples[:, None, :] - tree[None, :, :], axis=-1)
            assert np.array_equal(idxs, np.argmin(d, axis=1)) and np.allclose(dists, d.min(1), rtol=1e-5)
            assert ss.shape == samples.shape
        else:
            d = np.linalg.norm(tree - samples[0], axis=-1)
            assert np.array_equal(idxs, np.argsort(d, kind="stable")[:batch]) and np.allclose(dists, np.sort(d)[:batch], rtol=1e-5)
            assert ss.shape == (batch, 7) and (ss == samples[0]).all()


def test_batch_rrt_plan_builds_a_consistent_tree(stack):
    """batch_RRT_plan / global_explore / SearchTree (batch_tsa.py:11-189, search_tree.py:6-92): result dict, tree
    invariants, and every stored edge state hard-collision-free"""
    env, robot, dyn, ctrl = stack("magician")
    from cbfrrt.motion_planning.baseline import batch_RRT_plan, SearchTree
    x = dyn.sample_state_space(200)
    xs = x[dyn.safe_mask(x)][:, :4].numpy()
    init = xs[0]
    goal = init + 0.25 * (xs[1] - init) / np.linalg.norm(xs[1] - init)     # close by: reachable within a few extensions
    dyn.set_goal(goal)
    res = batch_RRT_plan(env, dyn, init, goal, RRT_PARAM=20, controller=ctrl, T=60, model_eps=0.3, steer_type="cbf",
                         batch=8, stop_when_success=True, rng=np.random.RandomState(0))
    assert set(res) >= {"success", "path", "edge_states", "nominal_path", "path_cost", "explored_nodes", "time_dict"}
    tree = res["search_tree"]
    n_nodes = tree.states.shape[0]
    assert n_nodes == len(tree.parents) == len(tree.x_list) == len(tree.in_goal_region) and n_nodes > 8
    assert all(p is None or 0 <= p < i for i, p in enumerate(tree.parents))
    assert tree.non_terminal_states.shape[0] == len(tree.non_terminal_idxes)
    assert np.array_equal(tree.non_terminal_states, tree.states[tree.non_terminal_idxes])
    edge_pts = np.array([pt for xl in tree.x_list[1:] for pt in xl])
    if len(edge_pts):
        assert not dyn.unsafe_mask(torch.as_tensor(edge_pts, dtype=torch.float32)).any()
    for i in range(1, n_nodes):     # an edge ends at its node -- for rows that stayed alive; a dead row's node is the
        if tree.x_list[i] and tree.freesp[i]:   # reference's running x_current (batch_tsa.py:231), beyond its stored edge
            assert np.allclose(tree.x_list[i][-1], tree.states[i], atol=1e-6)
    if res["success"]:
        assert np.allclose(res["path"][0], init) and np.linalg.norm(res["path"][-1] - goal) <= 0.2 + 1e-6
        assert len(res["path"]) == len(res["path_cost"]) == len(res["edge_states"]) and res["path_cost"][-1] == 0
    # the tree's own path() contract (search_tree.py:24-47)
    t = SearchTree(root=np.zeros(4))
    assert t.path() == ([], [], [], [])


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_dynamics_mirrors_match_reference_python(stack, golden, robot):
    """closed_loop_dynamics without observation refresh (arm_dynamics.py:474-523) and out_of_bounds_mask
    (control_affine_system.py:200-216) against the outputs of the reference's own code (tests/golden/make_golden.py)"""
    env, rb, dyn, ctrl = stack(robot)
    g = golden[robot]
    x = torch.tensor(g["datax"])
    x_next = dyn.closed_loop_dynamics(x, torch.tensor(g["shared_u"]).float(), update_observation=False)
    assert x_next.shape == g["cld_x_next"].shape and torch.equal(x_next, torch.tensor(g["cld_x_next"]))
    assert torch.equal(dyn.out_of_bounds_mask(torch.tensor(g["oob_x"])), torch.tensor(g["oob_mask"]))
    up, lo = dyn.state_limits
    assert np.allclose(up.numpy(), g["x_upper"]) and np.allclose(lo.numpy(), g["x_lower"])


def test_basic_robot_kinematics(stack, models):
    env, robot, dyn, ctrl = stack("panda")
    from oracle import oracle_f64 as o
    q = np.array(robot.q0)
    links = robot.forward_kinematics([0, 6, 7, 8, 11], q)
    fr = o.fk(models["panda"], q)
    for (p, R), l in zip(links, [0, 6, 7, 8, 11]):
        assert np.allclose(p, fr[l][:3, 3], atol=2e-6) and np.allclose(R, fr[l][:3, :3], atol=2e-6)
    assert np.allclose(robot.get_joint_position(robot.body_joints), q)
    # Jacobian vs finite differences of the float64 FK
    J_lin, J_ang = robot.get_jacobian(list(q), 7, [0.0, 0.0, 0.05])
    assert J_lin.shape == (3, 7) and J_ang.shape == (3, 7)
    def point(qq):
        f = o.fk(models["panda"], qq)[7]
        return f[:3, :3] @ np.array([0, 0, 0.05]) + f[:3, 3]
    for j in range(7):
        d = np.zeros(7); d[j] = 1e-6
        assert np.allclose((point(q + d) - point(q - d)) / 2e-6, J_lin[:, j], atol=1e-5)
    assert isinstance(robot.check_self_collision_free(), bool)
