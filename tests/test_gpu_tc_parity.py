"""GPU suite (-m gpu), tensor-core path: every tcgen05 instantiation of the barrier-network kernels against the C
oracle, at batch sizes that take that path by the library's own dispatch (B > 16 384 rows) and, for small batches, by
forcing it through cbfrrt_set_mlp_path (include/cbfrrt.h).

Instantiations (csrc/kernels.cuh) and the test that names each:
    filter_kernel_tc<7,2>, filter_kernel_tc<4,2>                      test_cbf_filter_tc_matches_oracle / _golden_tiled
    safety_kernel_tc<Panda,2,false>   (8 boxes, brute force + culling) test_safety_filter_tc_matches_oracle[panda-8box]
    safety_kernel_tc<Panda,2,true>    (256 boxes, grid broad phase)    test_safety_filter_tc_matches_oracle[panda-256box-grid]
    safety_kernel_tc<Magician,2,false> (16 boxes)                      test_safety_filter_tc_matches_oracle[magician-16box]
    safety_kernel_tc<Magician,2,true> (40 boxes + 24 spheres, grid)    test_safety_filter_tc_matches_oracle[magician-64obs-grid]
    rollout_kernel_tc<Panda,2,false>  (BASELINE config 4 shape: 200 steps, control every 4th)  test_rollout_tc_config4_shape[panda-8box-*]
    rollout_kernel_tc<Panda,2,true>                                    test_rollout_tc_config4_shape[panda-64obs-grid-*]
    rollout_kernel_tc<Magician,2,true>                                 test_rollout_tc_config4_shape[magician-64obs-grid-*]
    rollout_kernel_tc<Magician,2,false>                                tests/test_gpu_parity.py::test_rollout_variants_tensor_core_path
Bars: masks / indices / o bit-exact, do/dq within 1e-5 (conftest.rel_err); h / J / u / r against the float64
end-to-end truth as stated in conftest.assert_filter_matches_truth (h on every row; J on every row off the ReLU kinks;
u within 1e-5 or the explicit QP conditioning bound, the fraction under the plain 1e-5 printed)."""
import numpy as np
import pytest
import torch

from conftest import golden_state_dict, rel_err, assert_filter_matches_truth

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"
B_TC = 20011          # > SMALL_BATCH (16 384) and not a multiple of the 128-row tile: the tail tile is partly idle


def _t(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


@pytest.fixture(autouse=True)
def _auto_path():
    from cbfrrt import ops
    ops.set_mlp_path("auto")
    yield
    ops.set_mlp_path("auto")


def _scene(boxes6, spheres, grid=True):
    from cbfrrt import ops
    boxes6 = np.zeros((0, 6), np.float32) if boxes6 is None else np.asarray(boxes6, np.float32).reshape(-1, 6)
    b8, s4 = ops.pack_scene(boxes6[:, :3], boxes6[:, 3:], spheres, device=DEV)
    gr = ops.build_grid(b8, s4) if (grid and b8.shape[0] + s4.shape[0] > 0) else ops.no_grid(DEV)
    return b8, s4, gr


def _workload(robot):
    from cbfrrt import workloads as W
    return W.config(0 if robot == "panda" else 1, scale=1e-9)   # constants + seeded network; the states are drawn below


def _net(w):
    from cbfrrt import ops
    return ops.pack_weights(w["state_dict"], w["n"], 48, 2, device=DEV)


def _cases():
    from cbfrrt import workloads as W
    return {
        "panda-8box": ("panda", W.random_boxes(8, 2, 0.05, 0.12), None),
        "panda-256box-grid": ("panda", W.random_boxes(256, 2, 0.01, 0.05), None),
        "magician-16box": ("magician", W.random_boxes(16, 2, 0.03, 0.12), None),
        "magician-64obs-grid": ("magician", W.random_boxes(40, 3, 0.02, 0.08), W.random_spheres(24, 5, 0.02, 0.08)),
        "panda-64obs-grid": ("panda", W.random_boxes(40, 3, 0.02, 0.08), W.random_spheres(24, 5, 0.02, 0.08)),
    }


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_cbf_filter_tc_matches_oracle(coracle, robot):
    """filter_kernel_tc<7,2> / <4,2> by the library's own dispatch (B > 16 384), tail tile included"""
    from cbfrrt import ops, workloads as W
    w = _workload(robot)
    n, B = w["n"], B_TC
    rng = np.random.default_rng(4)
    datax = np.concatenate([W.sample_states(robot, B, 2), rng.uniform(-0.1, 0.3, (B, 1)).astype(np.float32),
                            (3 * rng.normal(size=(B, n))).astype(np.float32)], 1)
    goal = W.sample_states(robot, B, 8)
    assert ops.get_mlp_path() == "auto"
    for goal_np, pen, lam in ((goal, 5000.0, 1.0), (goal[:1], 0.5, 1.0), (goal, 50.0, 0.1)):
        h, J, u, r = ops.cbf_filter(_t(datax), _t(goal_np), _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), lam, pen,
                                    48, 2, torch.empty(0, device=DEV))
        assert_filter_matches_truth(coracle, n, w["weights"], datax, goal_np, w["K"], w["u_lo"], w["u_hi"], lam, pen, h, J, u, r,
                                    name=f"filter_kernel_tc<{n}> pen={pen} lam={lam}")
    # the FP32-FMA kernels forced at the same size agree with the tensor-core kernels to tolerance but not bitwise:
    # two different kernels did run
    ops.set_mlp_path("fma")
    h2, J2, u2, r2 = ops.cbf_filter(_t(datax), _t(goal), _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), 0.1, 50.0, 48, 2,
                                    torch.empty(0, device=DEV))
    assert rel_err(h2.cpu().numpy(), h.cpu().numpy(), floor=0.1) < TOL and not torch.equal(h2, h)
    assert_filter_matches_truth(coracle, n, w["weights"], datax, goal, w["K"], w["u_lo"], w["u_hi"], 0.1, 50.0, h2, J2, u2, r2,
                                name=f"filter_kernel<{n}> (FP32-FMA) at the same size")
    # caller-supplied u_ref on the tensor-core path
    ops.set_mlp_path("auto")
    uref = rng.uniform(-1, 1, (B, n)).astype(np.float32)
    h, J, u, r = ops.cbf_filter(_t(datax), _t(goal[:1]), _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), 1.0, 5000.0,
                                48, 2, _t(uref))
    assert_filter_matches_truth(coracle, n, w["weights"], datax, goal[:1], w["K"], w["u_lo"], w["u_hi"], 1.0, 5000.0, h, J, u, r,
                                u_ref=uref, name=f"filter_kernel_tc<{n}> caller u_ref")


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_cbf_filter_tc_golden_tiled(golden, robot):
    """the vectors recorded from the reference's own Python (h, autograd Jacobian, u, r), tiled past 16 384 rows so that
    the tensor-core kernels answer them; and the same rows with the path forced at the fixture's own size"""
    from cbfrrt import ops
    g = golden[robot]
    n = g["datax"].shape[1] // 2
    rows = g["datax"].shape[0]
    reps = 16384 // rows + 3
    Wt = ops.pack_weights(golden_state_dict(g), n, 48, 2, device=DEV)
    tile = lambda a: np.tile(a, (reps,) + (1,) * (a.ndim - 1))
    for path, rep in (("auto", True), ("tc", False)):
        ops.set_mlp_path(path)
        T = tile if rep else (lambda a: a)
        for tag in ("shared", "rows"):
            goal = g[f"{tag}_goal"] if tag == "shared" else T(g[f"{tag}_goal"])
            h, J, u, r = ops.cbf_filter(_t(T(g["datax"])), _t(goal), Wt, _t(g["K"]), _t(g["u_lower"]), _t(g["u_upper"]),
                                        float(g["lam"]), float(g["penalty"]), 48, 2, torch.empty(0, device=DEV))
            assert h.shape[0] == (rows * reps if rep else rows)
            assert rel_err(h.cpu().numpy(), T(g[f"{tag}_h"][:, 0]), floor=0.1) < TOL
            assert rel_err(J.cpu().numpy(), T(g[f"{tag}_J"][:, 0]), floor=0.1) < TOL
            assert rel_err(u.cpu().numpy(), T(g[f"{tag}_u"])) < TOL
            assert np.abs(r.cpu().numpy() - T(g[f"{tag}_r"][:, 0])).max() < TOL


@pytest.mark.parametrize("case", ["panda-8box", "panda-256box-grid", "magician-16box", "magician-64obs-grid"])
def test_safety_filter_tc_matches_oracle(coracle, case):
    """fused q -> (masks, datax, h, J, u, r) on the tcgen05 kernels: brute-force / culled sweep (4-group CTA) and grid
    broad phase (3-group CTA), both robots, against the brute-force oracle"""
    from cbfrrt import ops, workloads as W
    robot, boxes, sph = _cases()[case]
    w = _workload(robot)
    n, rid, B = w["n"], ops.ROBOT_ID[robot], B_TC
    q = W.sample_states(robot, B, 23)
    goal = W.sample_states(robot, B, 24)
    b8, s4, gr = _scene(boxes, sph)
    for goal_np in (goal, goal[:1]):
        args = (_t(q), _t(goal_np), b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"],
                w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
        safe, unsafe, datax, h, J, u, r = ops.safety_filter(*args)
        o = coracle.safety_filter(robot, q, boxes, sph, True, w["weights"], 48, 2, goal_np, w["K"], w["u_lo"], w["u_hi"],
                                  w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
        assert np.array_equal(safe.cpu().numpy(), o["safe"]) and np.array_equal(unsafe.cpu().numpy(), o["unsafe"])
        assert np.array_equal(datax.cpu().numpy()[:, :n + 1], o["datax"][:, :n + 1])       # q, o: bit-exact
        assert rel_err(datax.cpu().numpy()[:, n + 1:], o["datax"][:, n + 1:]) < TOL         # do/dq
        # the filter stage against the float64 truth evaluated on the kernel's own observation row (q, o bit-exact above)
        assert_filter_matches_truth(coracle, n, w["weights"], datax.cpu().numpy(), goal_np, w["K"], w["u_lo"], w["u_hi"], w["lam"],
                                    w["penalty"], h, J, u, r, name=f"safety_kernel_tc {case} goal_rows={goal_np.shape[0]}")
        v2, u2 = ops.valid_control(*args)
        assert torch.equal(v2, safe) and torch.equal(u2, u)
    assert 0 < o["unsafe"].mean() < 1
    # the same rows on the FP32-FMA kernels: same booleans and observation bit for bit, h within tolerance but from a
    # different kernel (not bitwise equal)
    ops.set_mlp_path("fma")
    s3, u3, d3, h3, J3, uu3, r3 = ops.safety_filter(*args)
    assert torch.equal(s3, safe) and torch.equal(u3, unsafe) and torch.equal(d3[:, :n + 1], datax[:, :n + 1])
    assert rel_err(h3.cpu().numpy(), h.cpu().numpy(), floor=0.1) < TOL and not torch.equal(h3, h)


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_forced_tc_path_small_batch(coracle, robot):
    """planner-sized batches (1, 37, 1000 rows) forced onto the tensor-core kernels: partly idle tiles, one CTA"""
    from cbfrrt import ops, workloads as W
    w = _workload(robot)
    n, rid = w["n"], ops.ROBOT_ID[robot]
    boxes = W.random_boxes(8, 2, 0.05, 0.12)
    b8, s4, gr = _scene(boxes, None)
    ops.set_mlp_path("tc")
    assert ops.get_mlp_path() == "tc"
    for B in (1, 37, 1000):
        q = W.sample_states(robot, B, 100 + B)
        args = (_t(q), _t(w["goal"]), b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"],
                w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
        safe, unsafe, datax, h, J, u, r = ops.safety_filter(*args)
        o = coracle.safety_filter(robot, q, boxes, None, True, w["weights"], 48, 2, w["goal"], w["K"], w["u_lo"], w["u_hi"],
                                  w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
        assert np.array_equal(safe.cpu().numpy(), o["safe"]) and np.array_equal(datax.cpu().numpy()[:, n], o["datax"][:, n])
        assert rel_err(h.cpu().numpy(), o["h"], floor=0.1) < TOL and rel_err(J.cpu().numpy(), o["J"], floor=0.1) < TOL
        assert rel_err(u.cpu().numpy(), o["u"]) < 1e-3


def _rollout_report(name, qf, alive, done, qo, ao, do_):
    same = (alive == ao) & (done == do_)
    err = np.abs(qf[same] - qo[same]).max(1) if same.any() else np.zeros(0)
    qs = {f"p{p}": float(np.quantile(err, p / 100)) for p in (50, 90, 99, 99.9)} if err.size else {}
    print(f"[rollout parity] {name}: rows {len(ao)}, same events {same.mean():.5f}, "
          f"|q_final - oracle| quantiles {qs}, max {err.max() if err.size else 0:.3e}, "
          f"frac<1e-5 {np.mean(err < 1e-5):.5f}, frac<1e-4 {np.mean(err < 1e-4):.5f}, frac<1e-3 {np.mean(err < 1e-3):.5f}")
    return same, err


@pytest.mark.parametrize("keep_going", [False, True])
@pytest.mark.parametrize("case", ["panda-8box", "panda-64obs-grid", "magician-64obs-grid"])
def test_rollout_tc_config4_shape(coracle, case, keep_going):
    """BASELINE config 4 shape on the tcgen05 rollout kernels: 32 768 trajectories x 200 steps, control every 4th step,
    per-row goals, safe starts, with (steer semantics) and without (keep_going) early stop, brute-force and grid scenes.

    Closed-loop tolerance, stated as measured (the distribution is printed): the discrete events (collision step) of a
    trajectory agree with the oracle's on >= 98 % of the rows; on those rows the final state agrees within 1e-3 on
    >= 98 % and within 0.3 rad everywhere -- 200 steps of a feedback loop carry the 1e-6-relative difference of the two
    network evaluations along, and an ill-conditioned QP row amplifies it.  Teacher-forced part,
    bit-exact: every stored state is not unsafe under the oracle."""
    from cbfrrt import ops, workloads as W
    robot, boxes, sph = _cases()[case]
    w = _workload(robot)
    n, rid, T, steps = w["n"], ops.ROBOT_ID[robot], 32768, 200
    cand = W.sample_states(robot, 8 * T, 51)
    safe0 = coracle.collide(robot, cand, boxes, sph, True)["safe"].astype(bool)
    pool = cand[safe0]
    assert pool.shape[0] > 500
    q0 = np.ascontiguousarray(np.resize(pool, (T, n)))
    goal = W.sample_states(robot, T, 52)
    b8, s4, gr = _scene(boxes, sph)
    args = (b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"],
            w["thr_hard"], rid, 48, 2)
    want_traj = not keep_going
    qf, alive, done, traj = ops.rollout(_t(q0), _t(goal), steps, 4, 1 / 120, 0, 10, 0.1, want_traj, keep_going, *args)
    qo, ao, do_, _ = coracle.rollout(robot, q0, goal, steps, 4, 1 / 120, boxes, sph, True, w["weights"], 48, 2, w["K"],
                                     w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"],
                                     keep_going=keep_going)
    same, err = _rollout_report(f"{case} keep_going={keep_going}", qf.cpu().numpy(), alive.cpu().numpy(), done.cpu().numpy(),
                                qo, ao, do_)
    assert same.mean() > 0.98, same.mean()
    assert np.mean(err < 1e-3) > 0.98 and err.max() < 0.3, (np.mean(err < 1e-3), err.max())
    assert (ao == 0).sum() > 50 and (ao == 1).sum() > 50
    if keep_going:
        assert (done.cpu().numpy() == steps).all()
    else:
        tr, dn = traj.cpu().numpy(), done.cpu().numpy()
        pick = np.random.default_rng(0).choice(T, 400, replace=False)
        flat = np.concatenate([tr[b, :dn[b]] for b in pick if dn[b] > 0])
        assert not coracle.collide(robot, flat, boxes, sph, True)["unsafe"].any()


def test_batched_steer_reference_semantics_gpu(coracle):
    """cbfrrt_rollout stuck_mode = 1 (batch_tsa.py:191-231) on both MLP paths against the oracle's restatement: stuck
    rows are reported as collisions, dead rows keep integrating"""
    from cbfrrt import ops, workloads as W
    robot, boxes, sph = _cases()["panda-8box"]
    w = _workload(robot)
    T, steps = 3000, 60
    cand = W.sample_states(robot, 8 * T, 61)
    pool = cand[coracle.collide(robot, cand, boxes, sph, True)["safe"].astype(bool)]
    q0 = np.ascontiguousarray(np.resize(pool, (T, 7)))
    goal = W.sample_states(robot, T, 62)
    goal[: T // 3] = q0[: T // 3] + 0.02           # targets next to the start: the rows stall and trip the stuck test
    b8, s4, gr = _scene(boxes, sph)
    args = (b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"],
            w["thr_hard"], 0, 48, 2)
    for path in ("fma", "tc"):
        ops.set_mlp_path(path)
        for keep_going in (True, False):
            qf, alive, done, traj = ops.rollout(_t(q0), _t(goal), steps, 4, 1 / 120, 9, 10, 3e-2, True, keep_going, *args, 1)
            qo, ao, do_, to = coracle.rollout(robot, q0, goal, steps, 4, 1 / 120, boxes, sph, True, w["weights"], 48, 2,
                                              w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"],
                                              w["penalty"], stuck_lag=9, stuck_after=10, stuck_eps=3e-2, want_traj=True,
                                              keep_going=keep_going, stuck_mode=1)
            same, err = _rollout_report(f"stuck_mode=1 {path} keep_going={keep_going}", qf.cpu().numpy(),
                                        alive.cpu().numpy(), done.cpu().numpy(), qo, ao, do_)
            assert same.mean() > 0.99 and np.mean(err < 1e-4) > 0.99 and err.max() < 5e-2
            stalled = (ao[: T // 3] == 0) & (do_[: T // 3] > 10)
            assert stalled.sum() > T // 6, "the stalled rows must die of the stuck test"
            tr = traj.cpu().numpy()
            for b in np.nonzero(same)[0][:40]:
                assert np.abs(tr[b, :do_[b]] - to[b, :do_[b]]).max() < 1e-4


@pytest.mark.parametrize("robot,hidden,layers", [("panda", 64, 2), ("panda", 48, 3), ("magician", 32, 1), ("panda", 128, 4),
                                                 ("magician", 17, 0), ("panda", 256, 2)])
def test_generic_network_shapes(coracle, robot, hidden, layers):
    """cbf_hidden_size / cbf_hidden_layers other than 48 x 2 (train_arm_mindis.py:176-177): the generic FP32 kernel behind
    cbfrrt_cbf_filter and cbfrrt_safety_filter against the float64 truth / the oracle, and through the controller mirror"""
    from cbfrrt import ops, workloads as W
    w = _workload(robot)
    n, rid, B = w["n"], ops.ROBOT_ID[robot], 3000
    sd, _ = W.make_state_dict(n, seed=5, hidden=hidden, layers=layers)
    wp = W.pack_weights_np(sd, n, hidden, layers)
    net = ops.pack_weights(sd, n, hidden, layers, device=DEV)
    boxes = W.random_boxes(8, 2, 0.05, 0.12)
    b8, s4, gr = _scene(boxes, None)
    q = W.sample_states(robot, B, 71)
    goal = W.sample_states(robot, B, 72)
    args = (_t(q), _t(goal), b8, s4, 1, gr, net, _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"],
            w["thr_hard"], rid, hidden, layers)
    safe, unsafe, datax, h, J, u, r = ops.safety_filter(*args)
    o = coracle.collide(robot, q, boxes, None, True, w["thr_soft"], w["thr_hard"])
    assert np.array_equal(safe.cpu().numpy(), o["safe"]) and np.array_equal(datax.cpu().numpy()[:, :n + 1], o["datax"][:, :n + 1])
    if hidden <= 128 and layers <= 4:          # the C oracle's own limits (MAXH, MAXL); the float64 truth shares them
        of = coracle.cbf_filter(n, wp, hidden, layers, datax.cpu().numpy(), goal, w["K"], w["u_lo"], w["u_hi"], w["lam"], w["penalty"])
        assert rel_err(h.cpu().numpy(), of["h"], floor=0.1) < TOL
        tr = coracle.cbf_filter_f64(n, wp, hidden, layers, datax.cpu().numpy(), goal, w["K"], w["u_lo"], w["u_hi"], w["lam"], w["penalty"])
        away = tr["margin"] > 1e-5
        assert away.mean() > 0.9
        assert rel_err(h.cpu().numpy(), tr["h"], floor=0.1) < TOL and rel_err(J.cpu().numpy()[away], tr["J"][away], floor=0.1) < TOL
        eu = np.abs(u.cpu().numpy() - tr["u"]).max(1)
        bound = TOL * np.maximum(1.0, np.abs(tr["u"]).max(1)) + 2.0 * (tr["kappa_c"] * np.abs(w["lam"] * (h.cpu().numpy() - tr["h"]))
                                                                      + tr["kappa_a"] * np.abs(J.cpu().numpy() - tr["J"]).max(1))
        assert (eu[away] <= bound[away]).all()
    else:                                      # beyond the oracle's table sizes: torch float64 on the CPU
        net64 = W.make_state_dict(n, seed=5, hidden=hidden, layers=layers)[1].double()
        x = torch.as_tensor(datax.cpu().numpy()[:, :n + 1], dtype=torch.float64)
        assert rel_err(h.cpu().numpy(), net64(x)[:, 0].detach().numpy(), floor=0.1) < TOL
    v2, u2 = ops.valid_control(*args)
    assert torch.equal(v2, safe) and torch.equal(u2, u)
    h2, J2, u3, r2 = ops.cbf_filter(datax, _t(goal), net, _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"], hidden, layers,
                                    torch.empty(0, device=DEV))
    assert torch.equal(h2, h) and torch.equal(u3, u)
    # the fused rollout is built for 48 x 2 only and says so
    from cbfrrt._lib import CbfrrtError
    with pytest.raises(CbfrrtError):
        ops.rollout(_t(q[:10]), _t(goal[:10]), 5, 4, 1 / 120, 0, 10, 0.1, False, False, b8, s4, 1, gr, net, _t(w["K"]), _t(w["u_lo"]),
                    _t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, hidden, layers)


@pytest.mark.parametrize("robot,T", [("panda", 1), ("panda", 50), ("magician", 128), ("panda", 7)])
def test_rollout_segments_matches_oracle(coracle, robot, T):
    """cbfrrt_rollout_segments: the nine 20-step segments of a planner expansion in one launch, incl. the batch-wide early
    exit (every row dead) and rows that stall, against the oracle's restatement of batch_tsa.py:154-170 / :191-231"""
    from cbfrrt import ops, workloads as W
    w = _workload(robot)
    n, rid = w["n"], ops.ROBOT_ID[robot]
    boxes = W.scene(3)[0]
    cand = W.sample_states(robot, 4000, 81 + T)
    pool = cand[coracle.collide(robot, cand, boxes, None, True)["safe"].astype(bool)]
    q0 = np.ascontiguousarray(pool[:T])
    goal = W.sample_states(robot, T, 82)
    goal[::3] = q0[::3] + 0.02                      # every third row stalls next to its start (stuck test)
    b8, s4, gr = _scene(boxes, None)
    n_seg, seg_len = 9, 20
    for use_grid in (True, False):
        g = gr if use_grid else ops.no_grid(DEV)
        sq, sa, sd, tr = ops.rollout_segments(_t(q0), _t(goal), n_seg, seg_len, 4, 1 / 120, 9, 10, 3e-2, b8, s4, 1, g, _net(w), _t(w["K"]),
                                              _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
        oq, oa, od, ot = coracle.rollout_segments(robot, q0, goal, n_seg, seg_len, 4, 1 / 120, boxes, None, True, w["weights"], 48, 2,
                                                  w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
        same = (sa.cpu().numpy() == oa) & (sd.cpu().numpy() == od)
        # a discrete event that differs in one segment moves everything after it in that row: compare rows up to there
        first_diff = np.where(same.all(1), n_seg, np.argmin(same, axis=1))
        assert (first_diff == n_seg).mean() >= 0.97, first_diff
        for b in range(T):
            k = first_diff[b]
            if k:
                assert np.abs(sq.cpu().numpy()[b, :k] - oq[b, :k]).max() < 2e-4
                for s_ in range(k):
                    m = od[b, s_]
                    assert m == 0 or np.abs(tr.cpu().numpy()[b, s_ * seg_len:s_ * seg_len + m] - ot[b, s_ * seg_len:s_ * seg_len + m]).max() < 2e-4
        assert (oa == 0).any() or T == 1
    from cbfrrt._lib import CbfrrtError
    with pytest.raises(CbfrrtError):               # more rows than one CTA holds
        ops.rollout_segments(_t(np.repeat(q0[:1], 129, 0)), _t(goal[:1]), 2, 20, 4, 1 / 120, 9, 10, 3e-2, b8, s4, 1, gr, _net(w),
                             _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)


@pytest.mark.parametrize("case", ["panda-256box-grid", "magician-16box", "panda-8box"])
def test_two_pass_sweep_equals_fused_kernel_and_oracle(coracle, case):
    """From 2^19 rows on the sweep runs as a geometry pass + a tensor-core barrier-network pass through datax (the caller's
    datax output, or the registered scratch): every output bit-identical to the single fused kernel, the fused exchange
    (peer stores from the second pass) included, and equal to the oracle on the first and last rows"""
    from cbfrrt import ops, workloads as W
    robot, boxes, sph = _cases()[case]
    w = _workload(robot)
    n, rid, B = w["n"], ops.ROBOT_ID[robot], (1 << 19) + 77
    q = ops.sample_states(B, 5, 0, rid, device=DEV)
    goal = _t(W.sample_states(robot, B, 24))
    b8, s4, gr = _scene(boxes, sph)
    res = {}
    try:
        for two_pass in (True, False):
            ops.set_two_pass(two_pass)
            for goal_t in (goal, goal[:1]):
                args = (q, goal_t, b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"],
                        w["thr_soft"], w["thr_hard"], rid, 48, 2)
                n0 = ops.launch_count()
                full = ops.safety_filter(*args)
                torch.cuda.synchronize()
                assert ops.launch_count() - n0 == (2 if two_pass else 1)
                n0 = ops.launch_count()
                v, u = ops.valid_control(*args)                       # no datax output: the registered scratch
                assert ops.launch_count() - n0 == (2 if two_pass else 1)
                assert torch.equal(v, full[0]) and torch.equal(u, full[5])
                # fused exchange into local "replicas" (one peer, plain pointers), at a row offset, ragged last tile
                G = B + 256
                vbuf, ubuf = torch.zeros(G, dtype=torch.uint8, device=DEV), torch.zeros((G, n), device=DEV)
                ops.valid_control_scatter(*args, [vbuf.data_ptr()], [ubuf.data_ptr()], 128, False)
                torch.cuda.synchronize()
                assert torch.equal(vbuf[128:128 + B], v) and torch.equal(ubuf[128:128 + B], u)
                assert not vbuf[:128].any() and not ubuf[128 + B:].any()
                res[(two_pass, goal_t.shape[0])] = full
    finally:
        ops.set_two_pass(True)
    for rows in (B, 1):
        for a, b in zip(res[(True, rows)], res[(False, rows)]):
            assert torch.equal(a, b)
    safe, unsafe, datax, h, J, u, r = res[(True, B)]
    for sl in (slice(0, 3000), slice(B - 3000, B)):
        qh, gh = q[sl].cpu().numpy(), goal[sl].cpu().numpy()
        o = coracle.safety_filter(robot, qh, boxes, sph, True, w["weights"], 48, 2, gh, w["K"], w["u_lo"], w["u_hi"],
                                  w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
        assert np.array_equal(safe[sl].cpu().numpy(), o["safe"]) and np.array_equal(unsafe[sl].cpu().numpy(), o["unsafe"])
        assert np.array_equal(datax[sl].cpu().numpy()[:, :n + 1], o["datax"][:, :n + 1])
        assert_filter_matches_truth(coracle, n, w["weights"], datax[sl].cpu().numpy(), gh, w["K"], w["u_lo"], w["u_hi"], w["lam"],
                                    w["penalty"], h[sl], J[sl], u[sl], r[sl], name=f"two-pass sweep {case}")
