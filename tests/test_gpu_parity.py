"""GPU suite (-m gpu): the CUDA path, called through the C-ABI (torch custom ops are thin ctypes wrappers over
include/cbfrrt.h), against the C oracle on the same seeded inputs.

Bars (BASELINE.json north_star): collision booleans and indices bit-exact; FK poses, barrier values, gradients
and QP controls within 1e-5 relative (conftest.rel_err: |a-b| / max(|b|, natural scale))."""
import numpy as np
import pytest
import torch

from conftest import golden_state_dict, rel_err, assert_controls_match

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"


def _t(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def _scene(boxes6, spheres, grid=True):
    """-> (boxes8, spheres4, grid): the broad-phase grid is ON by default, so every comparison below against the
    brute-force oracle also proves that the grid changes nothing."""
    from cbfrrt import ops
    boxes6 = np.zeros((0, 6), np.float32) if boxes6 is None else np.asarray(boxes6, np.float32).reshape(-1, 6)
    b8, s4 = ops.pack_scene(boxes6[:, :3], boxes6[:, 3:], spheres, device=DEV)
    gr = ops.build_grid(b8, s4) if (grid and b8.shape[0] + s4.shape[0] > 0) else ops.no_grid(DEV)
    return b8, s4, gr


def _net(w):
    from cbfrrt import ops
    sd = w["state_dict"]
    return ops.pack_weights(sd, w["n"], w["hidden"], w["layers"], device=DEV)


SCENES = {
    "default3": (np.array([[0.3, 0.17, 0.3, 0.05, 0.05, 0.1], [-0.4, 0.23, 0.4, 0.05, 0.05, 0.1],
                           [0.0, -0.23, 0.6, 0.05, 0.05, 0.1]], np.float32), None, True),
    "mixed": (None, None, True),  # filled in below from workloads
    "empty_floor": (None, None, True),
    "empty_nofloor": (None, None, False),
    "spheres_only_nofloor": (None, np.array([[0.3, -0.3, 0.5, 0.1], [0.0, 0.4, 0.7, 0.15]], np.float32), False),
}


def _get_scene(name):
    from cbfrrt import workloads as W
    if name == "mixed":
        return W.random_boxes(21, 2, 0.02, 0.1), W.random_spheres(13, 5, 0.02, 0.1), True
    return SCENES[name]


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_fk_matches_oracle(coracle, robot):
    from cbfrrt import ops, workloads as W
    q = W.sample_states(robot, 5000, 3)
    pos, rot = ops.fk(_t(q), ops.ROBOT_ID[robot])
    pos_o, rot_o = coracle.fk(robot, q)
    assert rel_err(pos.cpu().numpy(), pos_o) < TOL and rel_err(rot.cpu().numpy(), rot_o) < TOL
    # the deterministic-math contract makes FK bit-identical, not merely close
    assert np.array_equal(pos.cpu().numpy(), pos_o) and np.array_equal(rot.cpu().numpy(), rot_o)


@pytest.mark.parametrize("robot", ["panda", "magician"])
@pytest.mark.parametrize("scene", list(SCENES))
@pytest.mark.parametrize("B", [1, 127, 4096])
def test_collide_bit_exact(coracle, robot, scene, B):
    from cbfrrt import ops, workloads as W
    boxes, sph, floor = _get_scene(scene)
    q = W.sample_states(robot, B, 17 + B)
    b8, s4, gr = _scene(boxes, sph)
    safe, unsafe, datax, arg_link, arg_obs, mins = ops.collide(_t(q), b8, s4, int(floor), gr, 0.02, 0.0, ops.ROBOT_ID[robot])
    o = coracle.collide(robot, q, boxes, sph, floor, 0.02, 0.0)
    n = q.shape[1]
    assert np.array_equal(safe.cpu().numpy(), o["safe"]) and np.array_equal(unsafe.cpu().numpy(), o["unsafe"])
    assert np.array_equal(arg_link.cpu().numpy(), o["arg_link"]) and np.array_equal(arg_obs.cpu().numpy(), o["arg_obs"])
    d = datax.cpu().numpy()
    assert np.array_equal(d[:, :n + 1], o["datax"][:, :n + 1])          # q and the observation o: bit-exact
    assert rel_err(d[:, n + 1:], o["datax"][:, n + 1:]) < TOL          # do/dq
    assert np.array_equal(mins.cpu().numpy(), o["mins"])
    # the mask-only and observation-only instantiations agree with the full one
    s2, u2 = ops.masks(_t(q), b8, s4, int(floor), gr, 0.02, 0.0, ops.ROBOT_ID[robot])
    s3, u3, d3 = ops.observe(_t(q), b8, s4, int(floor), gr, 0.02, 0.0, ops.ROBOT_ID[robot])
    assert torch.equal(s2, safe) and torch.equal(u2, unsafe) and torch.equal(s3, safe) and torch.equal(d3, datax)


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_grid_broadphase_is_exact(robot):
    """every scene-consuming op gives bit-identical outputs with and without the nearest-candidate grid, incl.
    dense scenes that overflow cells and states whose spheres leave the grid region"""
    from cbfrrt import ops, workloads as W
    rid = ops.ROBOT_ID[robot]
    q = W.sample_states(robot, 20000, 4)
    q[:50] *= 3.0        # far outside the joint limits: still a valid FK input, spheres anywhere
    for boxes, sph in ((W.random_boxes(256, 2, 0.01, 0.05), None), (W.random_boxes(40, 3, 0.05, 0.3), W.random_spheres(24, 5, 0.05, 0.3)),
                       (np.array([[0.0, 0.0, 2.5, 0.2, 0.2, 0.2]], np.float32), None)):
        b8, s4, gr = _scene(boxes, sph)
        a = ops.collide(_t(q), b8, s4, 1, gr, 0.02, 0.0, rid)
        b = ops.collide(_t(q), b8, s4, 1, ops.no_grid(DEV), 0.02, 0.0, rid)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        qf, qt = _t(q[:4000]), _t(q[4000:8000])
        for x, y in zip(ops.edge_validity(qf, qt, 7, b8, s4, 1, gr, 0.02, 0.0, rid),
                        ops.edge_validity(qf, qt, 7, b8, s4, 1, ops.no_grid(DEV), 0.02, 0.0, rid)):
            assert torch.equal(x, y)


def test_collide_thresholds_and_strided_rows(coracle):
    from cbfrrt import ops, workloads as W
    boxes, sph, floor = _get_scene("mixed")
    q = W.sample_states("panda", 3000, 1)
    b8, s4, gr = _scene(boxes, sph)
    wide = torch.zeros(3000, 15, device=DEV)
    wide[:, :7] = _t(q)
    for ts, th in ((0.05, 0.0), (0.02, 0.02), (0.0, -0.01)):
        s, u = ops.masks(wide[:, :7], b8, s4, 1, gr, ts, th, 0)   # x[:, :q_dims] view: row stride 15, no copy
        o = coracle.collide("panda", q, boxes, sph, True, ts, th)
        assert np.array_equal(s.cpu().numpy(), o["safe"]) and np.array_equal(u.cpu().numpy(), o["unsafe"])


def test_collide_many_obstacles_and_limit(coracle):
    from cbfrrt import ops, workloads as W
    from cbfrrt._lib import CbfrrtError
    q = W.sample_states("panda", 2048, 9)
    boxes, sph = W.random_boxes(700, 2, 0.01, 0.04), W.random_spheres(324, 5, 0.01, 0.04)   # 1024 = the maximum
    b8, s4, gr = _scene(boxes, sph)
    safe, unsafe, datax, al, ao, mins = ops.collide(_t(q), b8, s4, 1, gr, 0.02, 0.0, 0)
    o = coracle.collide("panda", q, boxes, sph, True, 0.02, 0.0)
    assert np.array_equal(safe.cpu().numpy(), o["safe"]) and np.array_equal(ao.cpu().numpy(), o["arg_obs"])
    assert np.array_equal(datax.cpu().numpy()[:, 7], o["datax"][:, 7])
    b8big, _, _ = _scene(W.random_boxes(1025, 2, 0.01, 0.02), None, grid=False)
    with pytest.raises(CbfrrtError):
        ops.masks(_t(q), b8big, s4[:0], 1, ops.no_grid(DEV), 0.02, 0.0, 0)
    with pytest.raises(CbfrrtError):
        ops.build_grid(b8big, s4[:0])


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_cbf_filter_matches_reference_python_golden(golden, robot):
    """CUDA h / J / u / r against the vectors produced by the reference's own Python."""
    from cbfrrt import ops
    g = golden[robot]
    n = g["datax"].shape[1] // 2
    W = ops.pack_weights(golden_state_dict(g), n, 48, 2, device=DEV)
    for tag in ("shared", "rows"):
        h, J, u, r = ops.cbf_filter(_t(g["datax"]), _t(g[f"{tag}_goal"]), W, _t(g["K"]), _t(g["u_lower"]),
                                    _t(g["u_upper"]), float(g["lam"]), float(g["penalty"]), 48, 2,
                                    torch.empty(0, device=DEV))
        assert rel_err(h.cpu().numpy(), g[f"{tag}_h"][:, 0], floor=0.1) < TOL
        assert rel_err(J.cpu().numpy(), g[f"{tag}_J"][:, 0], floor=0.1) < TOL
        assert rel_err(u.cpu().numpy(), g[f"{tag}_u"]) < TOL
        assert np.abs(r.cpu().numpy() - g[f"{tag}_r"][:, 0]).max() < TOL
    h, J, u, r = ops.cbf_filter(_t(g["datax"]), _t(g["shared_goal"]), W, _t(g["K"]), _t(g["u_lower"]), _t(g["u_upper"]),
                                float(g["lam"]), 0.05, 48, 2, torch.empty(0, device=DEV))
    assert rel_err(u.cpu().numpy(), g["shared_u_pen0.05"]) < TOL
    assert rel_err(r.cpu().numpy(), g["shared_r_pen0.05"][:, 0], floor=0.1) < TOL


@pytest.mark.parametrize("robot,B", [("panda", 4096), ("magician", 10007)])
def test_cbf_filter_matches_oracle(coracle, robot, B):
    from cbfrrt import ops, workloads as W
    w = W.config(0 if robot == "panda" else 1, scale=0.001)
    n = w["n"]
    rng = np.random.default_rng(4)
    datax = np.concatenate([W.sample_states(robot, B, 2), rng.uniform(-0.1, 0.3, (B, 1)).astype(np.float32),
                            (3 * rng.normal(size=(B, n))).astype(np.float32)], 1)
    goal = W.sample_states(robot, B, 8)
    for goal_np, pen, lam in ((goal, 5000.0, 1.0), (goal[:1], 0.5, 1.0), (goal, 50.0, 0.1)):
        h, J, u, r = ops.cbf_filter(_t(datax), _t(goal_np), _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), lam, pen,
                                    48, 2, torch.empty(0, device=DEV))
        o = coracle.cbf_filter(n, w["weights"], 48, 2, datax, goal_np, w["K"], w["u_lo"], w["u_hi"], lam, pen)
        assert rel_err(h.cpu().numpy(), o["h"], floor=0.1) < TOL and rel_err(J.cpu().numpy(), o["J"], floor=0.1) < TOL
        uref_n = coracle.u_nominal(datax[:, :n], goal_np, w["K"], w["u_lo"], w["u_hi"])
        assert_controls_match(coracle, u, r, h, J, o, uref_n, w["u_lo"], w["u_hi"], lam, pen, frac=0.97)
        assert (np.abs(o["u"] - np.clip(-(datax[:, :n] - goal_np) @ w["K"].T, w["u_lo"], w["u_hi"])).max(1) > 1e-4).sum() > B // 20
    # caller-supplied u_ref (solve_CLF_QP(u_ref=...))
    uref = rng.uniform(-1, 1, (B, n)).astype(np.float32)
    h, J, u, r = ops.cbf_filter(_t(datax), _t(goal[:1]), _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), 1.0, 5000.0,
                                48, 2, _t(uref))
    o = coracle.cbf_filter(n, w["weights"], 48, 2, datax, goal[:1], w["K"], w["u_lo"], w["u_hi"], 1.0, 5000.0, u_ref=uref)
    assert_controls_match(coracle, u, r, h, J, o, uref, w["u_lo"], w["u_hi"], 1.0, 5000.0, frac=0.97)


def test_qp_op_matches_oracle_semantics(coracle):
    """cbfrrt::qp (generic CLFController path) on the QP data of the fused filter gives the same controls."""
    from cbfrrt import ops, workloads as W
    w = W.config(0, scale=0.001)
    B, n = 5000, 7
    rng = np.random.default_rng(6)
    datax = np.concatenate([W.sample_states("panda", B, 2), rng.uniform(-0.1, 0.3, (B, 1)).astype(np.float32),
                            (3 * rng.normal(size=(B, n))).astype(np.float32)], 1)
    o = coracle.cbf_filter(n, w["weights"], 48, 2, datax, w["goal"], w["K"], w["u_lo"], w["u_hi"], 1.0, 50.0)
    uref = np.clip(-(datax[:, :n] - w["goal"]) @ w["K"].T, w["u_lo"], w["u_hi"]).astype(np.float32)
    u, r = ops.qp(_t(o["J"]), _t(1.0 * o["h"]), _t(uref), _t(w["u_lo"]), _t(w["u_hi"]), 50.0)
    assert rel_err(u.cpu().numpy(), o["u"]) < 2e-5 and rel_err(r.cpu().numpy(), o["r"], floor=0.1) < 2e-5


def test_qp_reference_outside_box(coracle):
    """u_ref outside the control limits (caller-supplied u_ref, clf_controller.py:497-505)"""
    from cbfrrt import ops
    rng = np.random.default_rng(12)
    B, n, p = 20000, 7, 1e3
    lo, hi = -np.ones(n, np.float32), np.ones(n, np.float32)
    a = rng.normal(size=(B, n)).astype(np.float32)
    c = rng.normal(size=B).astype(np.float32)
    ur = rng.uniform(-3, 3, size=(B, n)).astype(np.float32)
    uo, ro = coracle.qp(a, c, ur, lo, hi, p)
    u, r = ops.qp(_t(a), _t(c), _t(ur), _t(lo), _t(hi), p)
    assert rel_err(u.cpu().numpy(), uo) < 2e-5 and rel_err(r.cpu().numpy(), ro, floor=0.1) < 2e-5


@pytest.mark.parametrize("n,p", [(7, 5000.0), (4, 5000.0), (7, 0.5), (1, 2.0), (2, 50.0), (3, 1e3), (5, 0.05), (6, 1e9),
                                 (8, 5000.0)])
def test_qp_backward_matches_oracle(coracle, n, p):
    """cbfrrt::qp_backward (closed-form reverse mode of the single-constraint QP) vs the fp64 oracle adjoint, which the
    CPU suite pins against finite differences (tests/test_qp_backward_cpu.py).  1e-5 relative to the row's gradient
    scale on every row; rows within 1e-6 of a kink of the solution map (where the two may pick different one-sided
    derivatives) are the only ones allowed to differ and must be rare."""
    from cbfrrt import ops
    rng = np.random.default_rng(200 + n)
    B = 50000
    lo, hi = -rng.uniform(0.5, 2.0, n).astype(np.float32), rng.uniform(0.5, 2.0, n).astype(np.float32)
    a = rng.normal(size=(B, n)).astype(np.float32)
    c = rng.normal(scale=2.0, size=B).astype(np.float32)
    ur = rng.uniform(-2.5, 2.5, size=(B, n)).astype(np.float32)
    gu = rng.normal(size=(B, n)).astype(np.float32)
    gr = rng.normal(size=B).astype(np.float32)
    ga_o, gc_o, gt_o, margin = coracle.qp_backward(a, c, ur, lo, hi, p, gu, gr, want_margin=True)
    ga, gc, gt = (x.cpu().numpy() for x in ops.qp_backward(_t(a), _t(c), _t(ur), _t(lo), _t(hi), p, _t(gu), _t(gr)))
    ok = margin > 1e-6
    assert ok.mean() > 0.999
    scale = np.maximum(1.0, np.maximum(np.abs(ga_o).max(1), np.maximum(np.abs(gt_o).max(1), np.abs(gc_o))))
    for got, want in ((ga, ga_o), (gt, gt_o), (gc[:, None], gc_o[:, None])):
        err = np.abs(got.astype(np.float64) - want).max(1) / scale
        assert err[ok].max() < TOL, (n, p, err[ok].max())
    # all three pieces of the solution map occur in the batch (or the two the penalty allows)
    u, r = coracle.qp(a, c, ur, lo, hi, p)
    g0 = c + np.sum(a * np.clip(ur, lo, hi), 1)
    assert (g0 <= 0).sum() > 100 and ((g0 > 0) & (r == 0)).sum() + (r > 0).sum() > 100
    # grad_r omitted = zeros; optional outputs; empty batch
    ga0, gc0, gt0 = ops.qp_backward(_t(a), _t(c), _t(ur), _t(lo), _t(hi), p, _t(gu), torch.empty(0, device="cuda"))
    ga_z, gc_z, gt_z = coracle.qp_backward(a, c, ur, lo, hi, p, gu, None)
    assert rel_err(ga0.cpu().numpy()[ok], ga_z[ok], floor=10.0) < TOL and rel_err(gc0.cpu().numpy()[ok], gc_z[ok], floor=10.0) < TOL
    e = ops.qp_backward(_t(a[:0]), _t(c[:0]), _t(ur[:0]), _t(lo), _t(hi), p, _t(gu[:0]), _t(gr[:0]))
    assert e[0].shape == (0, n) and e[1].shape == (0,)


@pytest.mark.parametrize("n,S,p", [(7, 2, 50.0), (7, 3, 1e3), (4, 2, 0.5), (7, 4, 0.05), (7, 2, 1e9), (8, 3, 1e3), (7, 1, 50.0),
                                   (5, 4, 50.0), (4, 3, 5000.0)])
def test_qp_multi_backward_matches_oracle(coracle, n, S, p):
    """cbfrrt::qp_multi_backward vs the fp64 oracle adjoint (pinned against finite differences of an independent KKT solve
    in the CPU suite).  S <= n: with more tight constraints than free controls the sensitivities are not unique."""
    from cbfrrt import ops
    rng = np.random.default_rng(300 + n * S)
    B = 20000
    f = lambda x: np.ascontiguousarray(x, np.float32)
    lo, hi = f(-rng.uniform(0.5, 2.0, n)), f(rng.uniform(0.5, 2.0, n))
    a, c, ur = f(rng.normal(size=(B, S, n))), f(rng.normal(scale=0.8, size=(B, S))), f(rng.uniform(-2.0, 2.0, size=(B, n)))
    gu, gr = f(rng.normal(size=(B, n))), f(rng.normal(size=(B, S)))
    ga_o, gc_o, gt_o, margin = coracle.qp_multi_backward(a, c, ur, lo, hi, p, gu, gr, want_margin=True)
    ga, gc, gt = (x.cpu().numpy() for x in ops.qp_multi_backward(_t(a), _t(c), _t(ur), _t(lo), _t(hi), p, 100, 1e-9, _t(gu), _t(gr)))
    ok = margin > 1e-5
    assert ok.mean() > 0.999
    scale = np.maximum(1.0, np.maximum(np.abs(ga_o).max((1, 2)), np.maximum(np.abs(gt_o).max(1), np.abs(gc_o).max(1))))
    err = np.maximum(np.abs(ga - ga_o).max((1, 2)), np.maximum(np.abs(gt - gt_o).max(1), np.abs(gc - gc_o).max(1))) / scale
    assert err[ok].max() < TOL, (n, S, p, err[ok].max())
    # autograd wrapper: same numbers through torch.autograd.Function
    A, Cc, T = (_t(x).requires_grad_(True) for x in (a[:256], c[:256], ur[:256]))
    u, r = ops.qp_multi_diff(A, Cc, T, _t(lo), _t(hi), p)
    ((u * _t(gu[:256])).sum() + (r * _t(gr[:256])).sum()).backward()
    assert np.array_equal(A.grad.cpu().numpy(), ga[:256]) and np.array_equal(T.grad.cpu().numpy(), gt[:256])
    e = ops.qp_multi_backward(_t(a[:0]), _t(c[:0]), _t(ur[:0]), _t(lo), _t(hi), p, 100, 1e-9, _t(gu[:0]), _t(gr[:0]))
    assert e[0].shape == (0, S, n)


def test_differentiable_filter_on_gpu(coracle):
    """solve_CLF_QP(requires_grad=True): torch autograd through h_nn for (h, J) + cbfrrt::qp / qp_backward.  Same
    controls as the fused inference launch; parameter gradients equal those of an all-torch float64 restatement
    (the QP piece differentiated by torch on the oracle's active set)."""
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    env = ArmEnv(["panda"], config_file="", include_floor=True)
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=env.robot_list[0])
    torch.manual_seed(1)
    ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0,
                                     safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48)
    with torch.no_grad():
        ctrl.h_nn.output_linear.bias += 30.0          # make the barrier constraint bite
    x = dyn.sample_safe(256)
    (u0, r0), _ = ctrl.solve_CLF_QP(x)
    (u1, r1), _ = ctrl.solve_CLF_QP(x, requires_grad=True)
    assert u1.requires_grad and rel_err(u1.detach().numpy(), u0.numpy()) < 2e-5
    assert (u0 - dyn.u_nominal(x)).abs().max() > 1e-3
    w = torch.linspace(-1, 1, 7)
    ((u1 * w).sum() + 0.01 * r1.sum()).backward()
    got = {k: p.grad.detach().cpu().double().clone() for k, p in ctrl.h_nn.named_parameters()}
    # float64 restatement: same network in double, QP adjoint from the oracle injected as the cotangent of (J, c)
    net64 = torch.nn.Sequential(*[torch.nn.Linear(m.in_features, m.out_features) if isinstance(m, torch.nn.Linear)
                                  else torch.nn.ReLU() for m in ctrl.h_nn]).double()
    with torch.no_grad():
        for (_, a_), (_, b_) in zip(ctrl.h_nn.state_dict().items(), net64.state_dict().items()):
            b_.copy_(a_.detach().cpu().double())
    d = x.double()
    xin = d[:, :8].clone().requires_grad_(True)
    h = net64(xin)
    Jh = torch.autograd.grad(h.sum(), xin, create_graph=True)[0]
    J = Jh[:, :7] + Jh[:, 7:] * d[:, 8:]
    up, lo_ = dyn.control_limits
    gu = np.tile(w.numpy(), (256, 1)).astype(np.float32)
    ga, gc, gt, margin = coracle.qp_backward(J.detach().numpy(), 0.1 * h.detach().numpy()[:, 0], dyn.u_nominal(x).numpy(),
                                             lo_.numpy(), up.numpy(), 5000.0, gu, np.full(256, 0.01, np.float32),
                                             want_margin=True)
    assert (margin > 1e-6).all()
    ((J * torch.from_numpy(ga).double()).sum() + (0.1 * h[:, 0] * torch.from_numpy(gc).double()).sum()).backward()
    for (k, g), (_, p64) in zip(got.items(), net64.named_parameters()):
        ref = p64.grad
        assert float((g - ref).abs().max()) < 1e-3 * max(1.0, float(ref.abs().max())), k


@pytest.mark.parametrize("n,S,p", [(4, 2, 50.0), (7, 3, 1e3), (7, 4, 0.05), (7, 2, 1e9), (4, 3, 5000.0), (7, 1, 50.0),
                                   (1, 2, 50.0), (2, 4, 5000.0), (8, 3, 1e3), (3, 2, 1e9), (5, 4, 50.0), (6, 2, 0.05)])
def test_qp_multi_matches_oracle(coracle, n, S, p):
    """cbfrrt::qp_multi (dual active-set solver) vs the fp64 oracle + a KKT certificate on a sample of rows"""
    from cbfrrt import ops
    from conftest import qp_kkt_residual
    rng = np.random.default_rng(100 + n * S)
    B = 20000
    lo, hi = -np.ones(n, np.float32), np.ones(n, np.float32)
    a = rng.normal(size=(B, S, n)).astype(np.float32)
    c = rng.normal(scale=0.8, size=(B, S)).astype(np.float32)
    ur = rng.uniform(-1.5, 1.5, size=(B, n)).astype(np.float32)
    uo, ro, ito = coracle.qp_multi(a, c, ur, lo, hi, p, max_iter=100, tol=1e-9, want_iters=True)
    u, r, iters = ops.qp_multi(_t(a), _t(c), _t(ur), _t(lo), _t(hi), p, 100, 1e-9)
    u, r, iters = u.cpu().numpy(), r.cpu().numpy(), iters.cpu().numpy()
    if S <= n:
        assert iters.max() <= 12, "active set not identified within a handful of iterations"
    else:       # more constraints than controls: a few rows per 10 000 zig-zag between active sets and report max_iter
        assert np.mean(iters <= 12) > 0.999 and np.mean(ito <= 12) > 0.999
        keep = (iters <= 12) & (ito <= 12)
        a, c, ur, u, r, uo, ro = a[keep], c[keep], ur[keep], u[keep], r[keep], uo[keep], ro[keep]
        B = len(u)
    pc = min(p, 1e6)
    assert rel_err(u, uo) < 1e-5
    obj = lambda uu, rr: ((uu.astype(np.float64) - ur) ** 2).sum(1) + pc * rr.astype(np.float64).sum(1)
    # p * r carries p * (fp32 rounding of the stored u in g ~ 1e-6) for constraints that are active but not relaxed
    assert np.all(obj(u, r) <= obj(uo, ro) * (1 + 1e-5) + 1e-5 + 5e-6 * pc)
    g = c + np.einsum("bsk,bk->bs", a, u)
    assert np.all(g - r <= 2e-5 * (1 + np.abs(c) + np.abs(a).sum(2))) and np.all(r >= 0)
    if pc <= 1e3:
        for b in range(0, B, 400):
            assert qp_kkt_residual(a[b], c[b], ur[b], lo, hi, pc, u[b], tol=1e-5) <= 1e-4 * max(1.0, pc * 1e-2), b
    if S == 1:
        u1, r1 = ops.qp(_t(a[:, 0]), _t(c[:, 0]), _t(ur), _t(lo), _t(hi), p)
        assert rel_err(u1.cpu().numpy(), u) < 1e-5 and np.allclose(r1.cpu().numpy(), r[:, 0], atol=1e-5)
        assert iters.max() <= 2


def test_qp_multi_infeasible_systems_relax(coracle):
    """S = n = 4 random constraints with the clamped 1e6 penalty: many rows are infeasible and must relax (mu = p);
    the multipliers reach 1e5..1e6 while free coordinates stay O(1) (the reason the kernel iterates in double)"""
    from cbfrrt import ops
    rng = np.random.default_rng(9)
    B, n, S, p = 20000, 4, 4, 1e9
    lo, hi = -np.ones(n, np.float32), np.ones(n, np.float32)
    a = rng.normal(size=(B, S, n)).astype(np.float32)
    c = rng.normal(scale=0.8, size=(B, S)).astype(np.float32)
    ur = rng.uniform(-1.5, 1.5, size=(B, n)).astype(np.float32)
    uo, ro, ito = coracle.qp_multi(a, c, ur, lo, hi, p, want_iters=True)
    u, r, iters = ops.qp_multi(_t(a), _t(c), _t(ur), _t(lo), _t(hi), p, 100, 1e-9)
    u, r, iters = u.cpu().numpy().astype(np.float64), r.cpu().numpy().astype(np.float64), iters.cpu().numpy()
    assert ito.max() <= 12 and iters.max() <= 12
    assert np.mean((ro > 1e-3).any(1)) > 0.1, "the case must actually exercise relaxed constraints"
    err = np.abs(u - uo).max(1)
    assert np.mean(err < 1e-5) > 0.999 and err.max() < 1e-3, (np.mean(err < 1e-5), err.max())
    obj = lambda uu, rr: ((uu - ur) ** 2).sum(1) + 1e6 * rr.sum(1)
    assert np.all(obj(u, r) <= obj(uo.astype(np.float64), ro.astype(np.float64)) * (1 + 1e-4) + 100.0)


@pytest.mark.parametrize("cfg,scale", [(0, 1.0), (1, 1 / 64)])
def test_safety_filter_fused_configs(coracle, cfg, scale):
    """BASELINE configs[0] in full (4096 Panda states, the reference's CPU-runnable case) and configs[1] at 1/64."""
    from cbfrrt import ops, workloads as W
    w = W.config(cfg, scale=scale)
    b8, s4, gr = _scene(w["boxes"], w["spheres"])
    rid = ops.ROBOT_ID[w["robot"]]
    args = (_t(w["q"]), _t(w["goal"]), b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"],
            w["penalty"], w["thr_soft"], w["thr_hard"], rid, 48, 2)
    safe, unsafe, datax, h, J, u, r = ops.safety_filter(*args)
    o = coracle.safety_filter(w["robot"], w["q"], w["boxes"], w["spheres"], True, w["weights"], 48, 2, w["goal"],
                              w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
    n = w["n"]
    assert np.array_equal(safe.cpu().numpy(), o["safe"]) and np.array_equal(unsafe.cpu().numpy(), o["unsafe"])
    assert np.array_equal(datax.cpu().numpy()[:, :n + 1], o["datax"][:, :n + 1])
    assert rel_err(datax.cpu().numpy()[:, n + 1:], o["datax"][:, n + 1:]) < TOL
    assert rel_err(h.cpu().numpy(), o["h"], floor=0.1) < TOL and rel_err(J.cpu().numpy(), o["J"], floor=0.1) < TOL
    uref_n = coracle.u_nominal(w["q"], w["goal"], w["K"], w["u_lo"], w["u_hi"])
    assert_controls_match(coracle, u, r, h, J, o, uref_n, w["u_lo"], w["u_hi"], w["lam"], w["penalty"])
    v2, u2 = ops.valid_control(*args)
    assert torch.equal(v2, safe) and torch.equal(u2, u)
    assert 0 < o["safe"].mean() < 1


@pytest.mark.parametrize("robot,K", [("panda", 32), ("panda", 5), ("magician", 32), ("panda", 200)])
def test_edge_validity_bit_exact(coracle, robot, K):
    from cbfrrt import ops, workloads as W
    w = W.config(2, scale=1 / 4096)
    E = 700 if K < 128 else 150
    qf = W.sample_states(robot, E, 0)
    m_lo, m_hi = np.array(W.ROBOT_MODELS[robot]["q_lower"], np.float32), np.array(W.ROBOT_MODELS[robot]["q_upper"], np.float32)
    qt = np.clip(qf + np.random.default_rng(1).uniform(-0.35, 0.35, qf.shape).astype(np.float32), m_lo, m_hi)
    b8, s4, gr = _scene(w["boxes"], w["spheres"])
    valid, first_bad = ops.edge_validity(_t(qf), _t(qt), K, b8, s4, 1, gr, 0.02, 0.0, ops.ROBOT_ID[robot])
    vo, fo = coracle.edge_validity(robot, qf, qt, K, w["boxes"], w["spheres"], True, 0.02, 0.0)
    assert np.array_equal(valid.cpu().numpy(), vo) and np.array_equal(first_bad.cpu().numpy(), fo)
    assert 0 < vo.mean() < 1


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_edge_validity_per_edge_lengths_bit_exact(coracle, robot):
    """cbfrrt::edge_validity_var: every edge with its own K_e = int(d/eps) (tsa.py:277-279), including K_e = 0"""
    from cbfrrt import ops, workloads as W
    E, n = 3000, (7 if robot == "panda" else 4)
    rng = np.random.default_rng(21)
    qf = W.sample_states(robot, E, 4)
    m = W.ROBOT_MODELS[robot]
    lo, hi = np.array(m["q_lower"], np.float32), np.array(m["q_upper"], np.float32)
    qt = np.clip(qf + rng.uniform(-0.4, 0.4, size=qf.shape).astype(np.float32) * rng.uniform(0, 1, size=(E, 1)).astype(np.float32), lo, hi)
    K = (np.linalg.norm(qf.astype(np.float64) - qt, axis=1) / 0.03).astype(np.int32)
    K[:5] = 0
    assert K.max() > 15 and (K == 0).sum() >= 5
    boxes, spheres = W.random_boxes(12, 3, 0.03, 0.1), W.random_spheres(6, 4, 0.03, 0.08)
    vo, fo = coracle.edge_validity_var(robot, qf, qt, K, boxes, spheres, True, 0.02, 0.0)
    b8, s4 = ops.pack_scene(boxes[:, :3], boxes[:, 3:], spheres, device="cuda")
    v, f = ops.edge_validity_var(_t(qf), _t(qt), torch.from_numpy(K), b8, s4, 1, ops.no_grid("cuda"), 0.02, 0.0,
                                 ops.ROBOT_ID[robot])
    assert np.array_equal(v.cpu().numpy(), vo) and np.array_equal(f.cpu().numpy(), fo)
    assert 0.05 < vo.mean() < 0.95
    # same K for all edges == the fixed-K entry point
    v2, f2 = ops.edge_validity(_t(qf), _t(qt), 9, b8, s4, 1, ops.no_grid("cuda"), 0.02, 0.0, ops.ROBOT_ID[robot])
    v3, f3 = ops.edge_validity_var(_t(qf), _t(qt), torch.full((E,), 9, dtype=torch.int32), b8, s4, 1, ops.no_grid("cuda"),
                                   0.02, 0.0, ops.ROBOT_ID[robot])
    assert torch.equal(v2, v3) and torch.equal(f2, f3)


@pytest.mark.parametrize("n,M,N,k", [(7, 50, 20000, 1), (4, 300, 5000, 1), (7, 1, 30000, 50), (7, 3, 40, 64), (3, 129, 513, 5)])
def test_knn_bit_exact(coracle, n, M, N, k):
    """cbfrrt::knn (planner front-end, batch_tsa.py:135-145) vs the oracle: indices and distances bit-identical"""
    from cbfrrt import ops
    rng = np.random.default_rng(n * 1000 + M)
    tree = rng.uniform(-2, 2, size=(N, n)).astype(np.float32)
    q = rng.uniform(-2, 2, size=(M, n)).astype(np.float32)
    tree[N // 2] = tree[3]
    q[0] = tree[3]                                  # exact tie between rows 3 and N//2
    io, do = coracle.knn(q, tree, k)
    i, d = ops.knn(_t(q), _t(tree), k)
    assert np.array_equal(i.cpu().numpy(), io) and np.array_equal(d.cpu().numpy(), do)
    assert io[0, 0] == 3
    w = ops.radius_mask(_t(q), _t(tree[:2000]), 1.5)
    assert np.array_equal(w.cpu().numpy(), coracle.radius_mask(q, tree[:2000], 1.5))


def test_edge_validity_equals_pointwise_safe_mask():
    """size-independent property: an edge is valid iff every interpolated configuration is safe (tsa.py:272-286)."""
    from cbfrrt import ops, workloads as W
    w = W.config(2, scale=1 / 1024)
    qf, qt, K = _t(w["q"]), _t(w["q_to"]), 32
    b8, s4, gr = _scene(w["boxes"], w["spheres"])
    valid, first_bad = ops.edge_validity(qf, qt, K, b8, s4, 1, gr, 0.02, 0.0, 0)
    bad = torch.full((qf.shape[0],), 1 << 30, device=DEV, dtype=torch.int64)
    for k in range(K + 1):
        # the interpolation contract is c = fmaf(k/K, q_to - q_from, q_from); float64 holds the product exactly
        t = np.float64(np.float32(k) / np.float32(K))
        d = (w["q_to"] - w["q"]).astype(np.float64)
        c = qt if k == K else _t((t * d + w["q"].astype(np.float64)).astype(np.float32))
        s, _ = ops.masks(c, b8, s4, 1, gr, 0.02, 0.0, 0)
        bad = torch.where((s == 0) & (bad > k), torch.full_like(bad, k), bad)
    exp_valid = (bad == (1 << 30))
    assert torch.equal(valid.bool(), exp_valid)
    assert torch.equal(first_bad.long()[~exp_valid], bad[~exp_valid])


def test_sampler_bit_exact_and_shard_invariant(coracle):
    from cbfrrt import ops
    for robot in ("panda", "magician"):
        rid = ops.ROBOT_ID[robot]
        a = ops.sample_states(100000, 7, 12345, rid).cpu().numpy()
        assert np.array_equal(a, coracle.sample_states(robot, 7, 12345, 100000))
        b = torch.cat([ops.sample_states(33333, 7, 12345, rid), ops.sample_states(66667, 7, 12345 + 33333, rid)])
        assert np.array_equal(a, b.cpu().numpy())
    big = ops.sample_states(10, 7, (1 << 33) + 5, 0).cpu().numpy()   # rows beyond 2^32
    assert np.array_equal(big, coracle.sample_states("panda", 7, (1 << 33) + 5, 10))


def test_rollout_against_oracle(coracle):
    """Closed loop (steer semantics).  The QP runs in fp32 on the GPU and fp64 in the oracle, so trajectories
    agree to tolerance, not bitwise; a discrete event (collision / stuck) may only differ where the oracle's own
    margin to that event is within tolerance."""
    from cbfrrt import ops, workloads as W
    w = W.config(3, scale=1 / 256)
    b8, s4, gr = _scene(w["boxes"], w["spheres"])
    T = w["q"].shape[0]
    safe0 = coracle.collide("panda", w["q"], w["boxes"], None, True)["safe"].astype(bool)
    q0, goal = w["q"][safe0][:600], w["goal"][safe0][:600]
    for steps, stuck in ((30, (9, 10, 0.1)), (60, (0, 10, 0.1))):
        qf, alive, done, traj = ops.rollout(_t(q0), _t(goal), steps, w["ctrl_every"], w["dt"], stuck[0], stuck[1], stuck[2],
                                            True, False, b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"],
                                            w["penalty"], w["thr_soft"], w["thr_hard"], 0, 48, 2)
        qo, ao, do_, to = coracle.rollout("panda", q0, goal, steps, w["ctrl_every"], w["dt"], w["boxes"], None, True,
                                          w["weights"], 48, 2, w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"],
                                          w["lam"], w["penalty"], stuck_lag=stuck[0], stuck_after=stuck[1],
                                          stuck_eps=stuck[2], want_traj=True)
        same = (alive.cpu().numpy() == ao) & (done.cpu().numpy() == do_)
        assert same.mean() > 0.99
        assert np.abs(qf.cpu().numpy()[same] - qo[same]).max() < 1e-4
        tr = traj.cpu().numpy()
        for b in np.nonzero(same)[0][:50]:
            assert np.abs(tr[b, :do_[b]] - to[b, :do_[b]]).max() < 1e-4
        # teacher-forced, bit-exact part: every appended state is not unsafe, and a dead trajectory's next state is
        dn = done.cpu().numpy()
        flat = np.concatenate([tr[b, :dn[b]] for b in range(len(dn)) if dn[b] > 0])
        assert not coracle.collide("panda", flat, w["boxes"], None, True)["unsafe"].any()
        assert (ao == 0).sum() > 0 and (do_ < steps).sum() > 0


def test_host_buffer_api_matches_device_api():
    from cbfrrt import ops, workloads as W
    w = W.config(1, scale=1 / 8)   # 131072 magician states: several chunks on 3 streams
    b8, s4, gr = _scene(w["boxes"], w["spheres"])
    B, n = w["q"].shape
    dev = ops.safety_filter(_t(w["q"]), _t(w["goal"]), b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]),
                            w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], 1, 48, 2)
    out = dict(safe=np.empty(B, np.uint8), unsafe=np.empty(B, np.uint8), datax=np.empty((B, 2 * n + 1), np.float32),
               h=np.empty(B, np.float32), J=np.empty((B, n), np.float32), u=np.empty((B, n), np.float32),
               r=np.empty(B, np.float32))
    ops.safety_filter_host("magician", w["q"], w["boxes"], w["spheres"], True, w["weights"], 48, 2, w["goal"], w["K"],
                           w["u_lo"], w["u_hi"], w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], out)
    for k, t in zip(("safe", "unsafe", "datax", "h", "J", "u", "r"), dev):
        assert np.array_equal(out[k], t.cpu().numpy()), k
    # per-row goals + a subset of outputs
    goal = W.sample_states("magician", B, 5)
    dev = ops.valid_control(_t(w["q"]), _t(goal), b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]),
                            w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], 1, 48, 2)
    out = dict(safe=np.empty(B, np.uint8), u=np.empty((B, n), np.float32))
    ops.safety_filter_host("magician", w["q"], w["boxes"], w["spheres"], True, w["weights"], 48, 2, goal, w["K"],
                           w["u_lo"], w["u_hi"], w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], out)
    assert np.array_equal(out["safe"], dev[0].cpu().numpy()) and np.array_equal(out["u"], dev[1].cpu().numpy())
    # the library keeps scene / network resident between calls: a changed scene (same sizes) and a changed network
    # must be re-uploaded, and going back must work too
    boxes2 = w["boxes"].copy()
    boxes2[:, 2] += 0.07
    sd2, _ = W.make_state_dict(n, seed=9)
    w2 = W.pack_weights_np(sd2, n)
    b8b, s4b, grb = _scene(boxes2, w["spheres"])
    net2 = ops.pack_weights(sd2, n, 48, 2, device="cuda")
    for bx, b8_, gr_, wt, net_ in ((boxes2, b8b, grb, w2, net2), (w["boxes"], b8, gr, w["weights"], _net(w)), (boxes2, b8b, grb, w["weights"], _net(w))):
        # (the broad phase belongs to ITS scene: the device-pointer API gets the grid built for the boxes it is given)
        dev = ops.valid_control(_t(w["q"]), _t(w["goal"]), b8_, s4, 1, gr_, net_, _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]),
                                w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], 1, 48, 2)
        ops.safety_filter_host("magician", w["q"], bx, w["spheres"], True, wt, 48, 2, w["goal"], w["K"],
                               w["u_lo"], w["u_hi"], w["lam"], w["penalty"], w["thr_soft"], w["thr_hard"], out)
        assert np.array_equal(out["safe"], dev[0].cpu().numpy()) and np.array_equal(out["u"], dev[1].cpu().numpy())


def test_full_size_config1_properties(coracle):
    """BASELINE configs[1] at full size (2^20 Magician states): split invariance (bitwise), oracle parity on a
    random 4096-row sample, and the QP feasibility / box constraints on every row."""
    from cbfrrt import ops, workloads as W
    w = W.config(1)
    b8, s4, gr = _scene(w["boxes"], w["spheres"])
    q = _t(w["q"])
    common = (_t(w["goal"]), b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"],
              w["thr_soft"], w["thr_hard"], 1, 48, 2)
    full = ops.safety_filter(q, *common)
    half = 333333
    a, b = ops.safety_filter(q[:half], *common), ops.safety_filter(q[half:], *common)
    for f, x, y in zip(full, a, b):
        assert torch.equal(f, torch.cat([x, y]))
    idx = np.random.default_rng(0).choice(q.shape[0], 4096, replace=False)
    o = coracle.safety_filter("magician", w["q"][idx], w["boxes"], w["spheres"], True, w["weights"], 48, 2, w["goal"],
                              w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"])
    assert np.array_equal(full[0].cpu().numpy()[idx], o["safe"]) and np.array_equal(full[1].cpu().numpy()[idx], o["unsafe"])
    safe, unsafe, datax, h, J, u, r = full
    uref_n = coracle.u_nominal(w["q"][idx], w["goal"], w["K"], w["u_lo"], w["u_hi"])
    assert_controls_match(coracle, u.cpu().numpy()[idx], r.cpu().numpy()[idx], h.cpu().numpy()[idx],
                          J.cpu().numpy()[idx], o, uref_n, w["u_lo"], w["u_hi"], w["lam"], w["penalty"])
    assert not (safe.bool() & unsafe.bool()).any()
    assert (u <= _t(w["u_hi"]) + 1e-6).all() and (u >= _t(w["u_lo"]) - 1e-6).all() and (r >= 0).all()
    g = w["lam"] * h + (J * u).sum(1) - r
    assert (g < 1e-4).all()


def test_python_api_drop_in_on_gpu(coracle):
    """The reference-shaped host API end to end: ArmEnv / ArmMindis / NeuralMindisCBFController / steer."""
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    from cbfrrt.motion_planning.baseline import steer, steer_fused, edge_checking, edge_checking_batch
    from cbfrrt import workloads as W
    env = ArmEnv(["panda"], config_file="", include_floor=True)
    robot = env.robot_list[0]
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=robot)
    torch.manual_seed(1)
    ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0,
                                     safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48,
                                     use_neural_actor=False)
    up, lo = dyn.state_limits
    assert (up > lo).all() and abs(dyn.K[0, 0].item() - 0.9834722117237553) < 1e-12  # reference value: B is float32 (golden K)
    x = dyn.sample_state_space(512)
    safe, unsafe = dyn.safe_mask(x), dyn.unsafe_mask(x)
    assert safe.dtype == torch.bool and safe.device == x.device and safe.shape == (512,)
    boxes6 = np.concatenate([env.obs_positions, env.obs_sizes], 1)
    o = coracle.collide("panda", x.numpy(), boxes6, None, True, 0.02, 0.0)
    assert np.array_equal(safe.numpy(), o["safe"].astype(bool)) and np.array_equal(unsafe.numpy(), o["unsafe"].astype(bool))
    assert torch.equal(dyn.boundary_mask(x), ~(safe | unsafe))
    datax = dyn.complete_sample_with_observations(x, 512)
    assert datax.shape == (512, 15) and np.array_equal(datax.numpy()[:, :8], o["datax"][:, :8])
    u, t_dict = ctrl.u(datax)
    assert u.shape == (512, 7) and isinstance(t_dict, dict) and {"nn_forward", "QP"} <= set(t_dict)
    h, J, extra = ctrl.h_with_jacobian(datax)
    assert h.shape == (512, 1) and J.shape == (512, 1, 7) and extra == {}
    net = ctrl.h_nn  # torch autograd of the same module = the reference's h_with_jacobian
    xg = datax[:, :8].clone().requires_grad_(True)
    hh = net(xg)
    Jh = torch.autograd.grad(hh.sum(), xg)[0]
    Jref = Jh[:, :7] + Jh[:, 7:] * datax[:, 8:]
    assert rel_err(h.numpy(), hh.detach().numpy(), floor=0.1) < TOL and rel_err(J[:, 0].numpy(), Jref.numpy(), floor=0.1) < TOL
    (u2, r2), _ = ctrl.solve_CLF_QP(datax, relaxation_penalty=0.5)
    assert u2.shape == (512, 7) and r2.shape == (512, 1)
    # steer: reference loop on our ops == one fused launch
    start = x[safe][0].numpy()
    target = x[safe][1].numpy()
    x_last, ok, td, x_list = steer(ctrl.u, dyn, target, start, RRT_step=30, device="cpu")
    x_last2, ok2, td2, x_list2 = steer_fused(ctrl, dyn, target, start, RRT_step=30)
    assert ok == ok2 and len(x_list) == len(x_list2) and np.abs(np.asarray(x_list[1:]) - np.asarray(x_list2[1:])).max() < 1e-4
    assert len(x_list) > 5
    # edge checking: reference loop == batched launch for the same K
    eps = 0.05
    K = int(np.linalg.norm(start - target) / eps)
    v, fb = edge_checking_batch(start[None], target[None], dyn, K)
    from cbfrrt.motion_planning.baseline import edge_checking_reference_loop
    assert bool(v.item()) == edge_checking(start, target, dyn, eps) == edge_checking_reference_loop(start, target, dyn, eps)
    assert robot.forward_kinematics([6, 7], torch.tensor(robot.q0))[1][0].shape == (3,)
    assert np.allclose(robot.forward_kinematics([7], robot.q0)[0][0], (0.288556, -0.257364, 0.361447), atol=2e-6)


def test_batch_rrt_plan_on_gpu():
    """the planner front-end end to end on the device ops: cbfrrt::knn -> cbfrrt::rollout -> cbfrrt::masks"""
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    from cbfrrt.motion_planning.baseline import batch_RRT_plan
    env = ArmEnv(["panda"], config_file="", include_floor=True)
    robot = env.robot_list[0]
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=robot)
    torch.manual_seed(1)
    ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0,
                                     safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48,
                                     use_neural_actor=False)
    x = dyn.sample_state_space(400)
    xs = x[dyn.safe_mask(x).cpu()][:, :7].cpu().numpy()
    init = xs[0]
    goal = init + 0.3 * (xs[1] - init) / np.linalg.norm(xs[1] - init)
    dyn.set_goal(goal)
    for steer_type, T in (("cbf", 100), ("line", 400)):
        res = batch_RRT_plan(env, dyn, init, goal, RRT_PARAM=20, controller=ctrl, T=T, model_eps=0.3, steer_type=steer_type,
                             batch=10, stop_when_success=True, rng=np.random.RandomState(0))
        tree = res["search_tree"]
        assert tree.states.shape[0] == len(tree.parents) > 1
        pts = np.array([pt[:7] for xl in tree.x_list[1:] for pt in xl])
        if len(pts):
            assert not dyn.unsafe_mask(torch.as_tensor(pts, dtype=torch.float32)).any()
        assert np.array_equal(tree.non_terminal_states, tree.states[tree.non_terminal_idxes])
        if res["success"]:
            assert np.allclose(res["path"][0], init) and np.linalg.norm(res["path"][-1][:7] - goal) <= 0.2 + 1e-5
    # the nominal (LQR) steer heads straight for the goal 0.3 rad away (0.05 rad per goal-biased extension): must be
    # found; the randomly initialised barrier network of this test may legitimately hold the CBF-filtered steer back
    assert res["success"]


def test_dataset_and_problem_generation_on_gpu(coracle, tmp_path):
    """SURVEY 8f rank 4 on the device ops: EpisodicDataModule.prepare_data (fixed quotas + noisy nominal rollouts) and
    propose_test_cases; every stored row / label is checked against the oracle in the scene it was drawn in."""
    import random
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.datamodules import EpisodicDataModule
    from cbfrrt.motion_planning.evaluation import propose_test_cases
    np.random.seed(5)
    env = ArmEnv(["panda"], config_file="", include_floor=True)
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.05, env=env, robot=env.robot_list[0])
    up, lo = dyn.state_limits
    dm = EpisodicDataModule(dyn, [(float(a), float(b)) for a, b in zip(lo, up)], max_episode=1, trajectories_per_episode=8,
                            trajectory_length=6, fixed_samples=1000, noise_level=0.1, batch_size=64,
                            quotas={"safe": 0.3, "unsafe": 0.3, "boundary": 0.2, "goal": 0.1}, data_root=str(tmp_path))
    random.seed(0), np.random.seed(0), torch.manual_seed(0)
    samples, masks = dm.sample_fixed()
    assert samples.shape == (900, 15) and masks["safe"][:300].all() and masks["unsafe"][300:600].all()
    boxes6 = np.concatenate([env.obs_positions, env.obs_sizes], 1)
    o = coracle.collide("panda", samples[:, :7].numpy(), boxes6, None, True, 0.05, 0.0)
    assert np.array_equal(masks["safe"].numpy(), o["safe"].astype(bool))
    assert np.array_equal(masks["unsafe"].numpy(), o["unsafe"].astype(bool))
    assert np.array_equal(samples.numpy()[:, :8], o["datax"][:, :8])
    assert rel_err(samples.numpy()[:, 8:], o["datax"][:, 8:], floor=1.0) < TOL
    x_sim, sim_masks = dm.sample_trajectories(dyn.noisy_simulator)     # resets the scene to a random training scene
    assert x_sim.shape == (8 * 6, 15)
    boxes6 = np.concatenate([env.obs_positions, env.obs_sizes], 1)
    o = coracle.collide("panda", x_sim[:, :7].numpy(), boxes6, None, True, 0.05, 0.0)
    assert np.array_equal(sim_masks["safe"].numpy(), o["safe"].astype(bool))
    assert np.array_equal(sim_masks["boundary"].numpy(), ~(o["safe"].astype(bool) | o["unsafe"].astype(bool)))
    assert np.array_equal(x_sim.numpy()[:, :8], o["datax"][:, :8])
    dm.prepare_data()
    data = torch.load(dm.dataset_path())
    assert data["x_training"].shape[0] + data["x_validation"].shape[0] == 900 + 48
    pos, siz, start, goal = propose_test_cases(dyn, obstacle_num=8)
    o = coracle.collide("panda", np.stack([start, goal]).astype(np.float32), np.concatenate([pos, siz], 1), None, True, 0.05, 0.0)
    assert o["safe"].all() and np.linalg.norm(start - goal) > 0.7


def test_registered_ops_equal_direct_calls():
    """torch.ops.cbfrrt.* (dispatcher) and the module-level direct entry points run the same implementation"""
    from cbfrrt import ops, workloads as W
    q = _t(W.sample_states("panda", 300, 3))
    a, b = ops.fk(q, 0), torch.ops.cbfrrt.fk(q, 0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert set(ops.registered) >= {"fk", "collide", "safety_filter", "rollout", "knn", "qp_multi", "edge_validity_var"}
    assert ops.fk.op is ops.registered["fk"]
    with pytest.raises(NotImplementedError):
        ops.fk(q.cpu(), 0)


def test_qp_multi_degenerate_constraints_gpu(coracle):
    """duplicate / zero / (anti-)parallel constraint rows on the device solver: same controls as the oracle"""
    from cbfrrt import ops
    from conftest import degenerate_qp_batch
    a, c, ur = degenerate_qp_batch()
    lo, hi = -np.ones(4, np.float32), np.ones(4, np.float32)
    for p in (50.0, 1e9):
        uo, ro = coracle.qp_multi(a, c, ur, lo, hi, p, max_iter=300)
        u, r, it = ops.qp_multi(_t(a), _t(c), _t(ur), _t(lo), _t(hi), p, 100, 1e-9)
        u, r = u.cpu().numpy(), r.cpu().numpy()
        assert np.isfinite(u).all() and np.isfinite(r).all()
        assert rel_err(u, uo) < 1e-5 and np.allclose(r, ro, atol=1e-4 + 1e-6 * min(p, 1e6))
        assert it.cpu().numpy().max() <= 12


def test_qp_single_edge_cases_gpu(coracle):
    """cbfrrt::qp on zero / partly zero / tiny / huge gradients, u_ref on the box, asymmetric limits"""
    from cbfrrt import ops
    from conftest import edge_case_qp_batch
    a, c, ur, lo, hi = edge_case_qp_batch()
    for p in (0.0, 0.05, 5000.0, 1e9):
        uo, ro = coracle.qp(a, c, ur, lo, hi, p)
        u, r = ops.qp(_t(a), _t(c), _t(ur), _t(lo), _t(hi), p)
        u, r = u.cpu().numpy(), r.cpu().numpy()
        assert np.isfinite(u).all() and np.isfinite(r).all() and (u >= lo).all() and (u <= hi).all() and (r >= 0).all()
        well = np.r_[0:900, 1100:len(c)]               # rows whose gradient is O(1): fp32 holds 2e-5
        assert rel_err(u[well], uo[well]) < 2e-5 and rel_err(r[well], ro[well], floor=0.1) < 2e-5
        # tiny / huge gradients: the answer is only as good as fp32 carries the constraint value c + a.u
        g = c[900:1100].astype(np.float64) + np.einsum("bk,bk->b", a[900:1100].astype(np.float64), u[900:1100])
        assert np.all(g - r[900:1100] <= 1e-5 * (1 + np.abs(c[900:1100]) + np.abs(a[900:1100]).sum(1)))


@pytest.mark.parametrize("robot", ["panda", "magician"])
@pytest.mark.parametrize("use_grid", [False, True])
def test_collide_degenerate_scene_elements(coracle, robot, use_grid):
    """zero-size boxes, a box that swallows part of the arm (sphere centres INSIDE: negative distances, interior
    normals), exact duplicates (arg-min ties -> lowest index), zero-radius spheres, obstacles outside the grid region"""
    from cbfrrt import ops, workloads as W
    rid, n = ops.ROBOT_ID[robot], (7 if robot == "panda" else 4)
    rng = np.random.default_rng(31)
    boxes = [[0.35, 0.0, 0.45, 0.25, 0.3, 0.3],          # big: the arm reaches into it
             [0.2, 0.3, 0.2, 0.0, 0.0, 0.0],              # a point
             [-0.3, 0.2, 0.6, 0.1, 0.0, 0.1],             # a plate
             [0.35, 0.0, 0.45, 0.25, 0.3, 0.3],           # exact duplicate of box 0
             [3.0, 3.0, 3.0, 0.1, 0.1, 0.1]]              # outside the grid region
    spheres = [[0.0, -0.4, 0.5, 0.0], [0.0, -0.4, 0.5, 0.0], [-0.4, -0.3, 0.3, 0.15], [5.0, 0.0, 0.0, 1.0]]
    if use_grid:                                          # the grid is only used from 40 obstacles on
        extra = W.random_boxes(40, 8, 0.01, 0.04)
        extra[:, :3] += np.array([0.0, 0.0, 0.9], np.float32)
        boxes = boxes + extra.tolist()
    boxes, spheres = np.array(boxes, np.float32), np.array(spheres, np.float32)
    q = W.sample_states(robot, 30000, 17)
    b8, s4 = ops.pack_scene(boxes[:, :3], boxes[:, 3:], spheres, device="cuda")
    gr = ops.build_grid(b8, s4) if use_grid else ops.no_grid("cuda")
    for floor in (1, 0):
        o = coracle.collide(robot, q, boxes, spheres, bool(floor), 0.02, 0.0)
        safe, unsafe, datax, arg_link, arg_obs, mins = ops.collide(_t(q), b8, s4, floor, gr, 0.02, 0.0, rid)
        assert np.array_equal(safe.cpu().numpy(), o["safe"]) and np.array_equal(unsafe.cpu().numpy(), o["unsafe"])
        assert np.array_equal(datax.cpu().numpy()[:, n], o["datax"][:, n])
        assert np.array_equal(arg_link.cpu().numpy(), o["arg_link"]) and np.array_equal(arg_obs.cpu().numpy(), o["arg_obs"])
        assert rel_err(datax.cpu().numpy()[:, n + 1:], o["datax"][:, n + 1:]) < 1e-5
        assert (o["datax"][:, n] < 0).mean() > 0.02, "the case must contain penetrating states"
        off = 1 if floor else 0
        assert not np.any(o["arg_obs"] == off + 3), "a duplicate never wins against its lower-index twin"


@pytest.mark.parametrize("keep_going", [False, True])
def test_rollout_variants_tensor_core_path(coracle, keep_going):
    """Magician, 20000 trajectories (tcgen05 kernels), 25 steps with ctrl_every = 3 (steps not a multiple), no stored
    trajectory, keep_going on / off, per-row goals; steps = 0 is a no-op"""
    from cbfrrt import ops, workloads as W
    w = W.config(1, scale=1e-9)
    b8, s4, gr = _scene(w["boxes"], w["spheres"])
    T = 20000
    q0 = W.sample_states("magician", T, 41)
    goal = W.sample_states("magician", T, 42)
    args = (b8, s4, 1, gr, _net(w), _t(w["K"]), _t(w["u_lo"]), _t(w["u_hi"]), w["lam"], w["penalty"], w["thr_soft"],
            w["thr_hard"], 1, 48, 2)
    qf, alive, done, traj = ops.rollout(_t(q0), _t(goal), 25, 3, 1 / 120, 0, 10, 0.1, False, keep_going, *args)
    assert traj.numel() == 0
    qo, ao, do_, _ = coracle.rollout("magician", q0, goal, 25, 3, 1 / 120, w["boxes"], None, True, w["weights"], 48, 2,
                                     w["K"], w["u_lo"], w["u_hi"], w["thr_soft"], w["thr_hard"], w["lam"], w["penalty"],
                                     keep_going=keep_going)
    same = (alive.cpu().numpy() == ao) & (done.cpu().numpy() == do_)
    assert same.mean() > 0.995
    # an ill-conditioned QP (a free axis with a tiny |Lg_i|) amplifies the 1e-7 difference of the two MLP evaluations
    # and the closed loop carries it along: 1e-4 on all but a few rows, bounded everywhere
    err = np.abs(qf.cpu().numpy()[same] - qo[same]).max(1)
    assert np.mean(err < 1e-4) > 0.995 and err.max() < 5e-2, (np.mean(err < 1e-4), err.max())
    assert (ao == 0).sum() > 100 and (ao == 1).sum() > 100
    if keep_going:
        assert (do_[ao == 0] == 25).mean() > 0.5 or (do_ <= 25).all()
    q1, a1, d1, _ = ops.rollout(_t(q0), _t(goal), 0, 3, 1 / 120, 0, 10, 0.1, False, keep_going, *args)
    assert torch.equal(q1.cpu(), torch.from_numpy(q0)) and (d1 == 0).all() and a1.all()


@pytest.mark.parametrize("cfg", [1, 4])
def test_full_size_obstacle_permutation_invariance(cfg):
    """BASELINE configs 1 (2^20 Magician states, 16 boxes, culled brute force) and 4 (2^21-state slice, 256 boxes,
    grid broad phase): minima and masks are bit-identical under a permutation of the obstacles, the arg-min follows it"""
    from cbfrrt import ops, workloads as W
    w = W.config(cfg, scale=1.0 if cfg == 1 else 1 / 32)
    rid, n = ops.ROBOT_ID[w["robot"]], w["n"]
    q = _t(w["q"]) if w["q"] is not None else ops.sample_states(w["n_states"], w["sample_seed"], 0, rid, device="cuda")
    boxes = w["boxes"]
    perm = np.random.default_rng(5).permutation(len(boxes))
    res = []
    for bx in (boxes, boxes[perm]):
        b8, s4 = ops.pack_scene(bx[:, :3], bx[:, 3:], w["spheres"], device="cuda")
        gr = ops.build_grid(b8, s4)
        res.append(ops.collide(q, b8, s4, 1, gr, 0.02, 0.0, rid))
    (s0, u0, d0, l0, o0, _), (s1, u1, d1, l1, o1, _) = res
    assert torch.equal(s0, s1) and torch.equal(u0, u1) and torch.equal(d0[:, n], d1[:, n]) and torch.equal(l0, l1)
    is_box = (o0 >= 1) & (o0 <= len(boxes))
    p = torch.as_tensor(perm, device="cuda")
    assert torch.equal(p[(o1[is_box] - 1).long()], (o0[is_box] - 1).long())
    assert is_box.float().mean() > 0.3 and u0.float().mean() > 0.02      # boxes do win the arg-min; collisions do occur


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_geometry_contract_fixture_gpu(robot):
    """the CUDA kernels against the committed geometry fixture (tests/golden/geometry_contract_*.npz): masks, minimum,
    arg-min and FK bit-identical, gradient within 1e-5"""
    import os
    from cbfrrt import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"geometry_contract_{robot}.npz"))
    rid, n = ops.ROBOT_ID[robot], g["q"].shape[1]
    b8, s4 = ops.pack_scene(g["boxes"][:, :3], g["boxes"][:, 3:], g["spheres"], device="cuda")
    for floor in (1, 0):
        safe, unsafe, datax, arg_link, arg_obs, _ = ops.collide(_t(g["q"]), b8, s4, floor, ops.no_grid("cuda"), 0.02, 0.0, rid)
        assert np.array_equal(safe.cpu().numpy(), g[f"safe_floor{floor}"]) and np.array_equal(unsafe.cpu().numpy(), g[f"unsafe_floor{floor}"])
        assert np.array_equal(datax.cpu().numpy()[:, :n + 1], g[f"datax_floor{floor}"][:, :n + 1])
        assert np.array_equal(arg_link.cpu().numpy(), g[f"arg_link_floor{floor}"]) and np.array_equal(arg_obs.cpu().numpy(), g[f"arg_obs_floor{floor}"])
        assert rel_err(datax.cpu().numpy()[:, n + 1:], g[f"datax_floor{floor}"][:, n + 1:]) < 1e-5
    pos, rot = ops.fk(_t(g["q"][:50]), rid)
    assert np.array_equal(pos.cpu().numpy(), g["fk_pos"]) and np.array_equal(rot.cpu().numpy().reshape(g["fk_rot"].shape), g["fk_rot"])


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_kernels_match_reference_systems_golden(robot):
    """the CUDA kernels against the outputs of the reference's own safe_mask / unsafe_mask /
    complete_sample_with_observations loops, recorded on the stand-in physics client (tests/golden/ref_systems_*.npz)"""
    from cbfrrt import ops
    from conftest import check_against_reference_systems_golden
    rid = ops.ROBOT_ID[robot]

    def collide(q, boxes, floor):
        b8, s4 = ops.pack_scene(boxes[:, :3], boxes[:, 3:], None, device="cuda")
        safe, unsafe, datax, _, _, mins = ops.collide(_t(q), b8, s4, int(floor), ops.no_grid("cuda"), 0.02, 0.0, rid)
        return dict(safe=safe.cpu().numpy(), unsafe=unsafe.cpu().numpy(), datax=datax.cpu().numpy(), mins=mins.cpu().numpy())

    check_against_reference_systems_golden(robot, collide)
