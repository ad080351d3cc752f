"""Reverse mode of the single-constraint CBF-QP (differentiable QP; reference: backpropagation
through the CvxpyLayer of clf_controller.py:82-139 via solve_CLF_QP(requires_grad=True) :461-528).

The oracle's adjoint (oracle/cbfrrt_oracle.c oracle_qp_backward, float64) is pinned here against central finite
differences of an INDEPENDENT float64 forward solve (bisection on the dual, numpy) on all three pieces of the solution
map, and the host-side autograd plumbing (ops.qp_diff, CLFController / NeuralMindisCBFController with
requires_grad=True) is exercised on the oracle-backed fakes.  The CUDA kernel is checked against the same oracle in
tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import fake_ops


def qp_forward64(a, c, t, lo, hi, p):
    """u(mu) = clip(t - mu a / 2); mu in [0, p] by bisection of the non-increasing g(mu) = c + a.u(mu) (float64)"""
    def u_of(mu):
        return np.clip(t - 0.5 * mu * a, lo, hi)
    g = lambda mu: c + a @ u_of(mu)
    if g(0.0) <= 0:
        return u_of(0.0), 0.0
    if g(p) > 0:
        return u_of(p), g(p)
    l, h = 0.0, p
    for _ in range(200):
        m = 0.5 * (l + h)
        if g(m) > 0:
            l = m
        else:
            h = m
    return u_of(0.5 * (l + h)), 0.0


def cases(n, B, seed, penalty):
    rng = np.random.default_rng(seed)
    lo, hi = -rng.uniform(0.5, 2.0, n), rng.uniform(0.5, 2.0, n)
    a = rng.normal(0, 1, (B, n))
    t = rng.uniform(-2.5, 2.5, (B, n))          # u_ref in and out of the box
    c = rng.normal(0, 2.0, B)
    f = lambda x: np.asarray(x, np.float32)
    return f(a), f(c), f(t), f(lo), f(hi), penalty


@pytest.mark.parametrize("n,penalty", [(7, 5000.0), (4, 5000.0), (7, 0.5), (1, 2.0), (8, 1e9)])
def test_oracle_adjoint_equals_finite_differences(coracle, n, penalty):
    B = 400
    a, c, t, lo, hi, p = cases(n, B, 10 + n, penalty)
    rng = np.random.default_rng(99)
    gu, gr = rng.normal(0, 1, (B, n)).astype(np.float32), rng.normal(0, 1, B).astype(np.float32)
    ga, gc, gt, margin = coracle.qp_backward(a, c, t, lo, hi, p, gu, gr, want_margin=True)
    pc = min(p, 1e6)
    a64, c64, t64, lo64, hi64 = (x.astype(np.float64) for x in (a, c, t, lo, hi))
    eps = 1e-6
    kinds = {"inactive": 0, "tight": 0, "relaxed": 0}
    checked = 0
    for b in range(B):
        if margin[b] < 1e-3:       # a kink within reach of the difference stencil: the map is not differentiable there
            continue
        u0, r0 = qp_forward64(a64[b], c64[b], t64[b], lo64, hi64, pc)
        g0 = c64[b] + a64[b] @ np.clip(t64[b], lo64, hi64)
        kinds["inactive" if g0 <= 0 else ("relaxed" if r0 > 0 else "tight")] += 1

        def loss(av, cv, tv):
            u, r = qp_forward64(av, cv, tv, lo64, hi64, pc)
            return float(gu[b].astype(np.float64) @ u + float(gr[b]) * r)
        fd_a, fd_t = np.zeros(n), np.zeros(n)
        for k in range(n):
            d = np.zeros(n); d[k] = eps
            fd_a[k] = (loss(a64[b] + d, c64[b], t64[b]) - loss(a64[b] - d, c64[b], t64[b])) / (2 * eps)
            fd_t[k] = (loss(a64[b], c64[b], t64[b] + d) - loss(a64[b], c64[b], t64[b] - d)) / (2 * eps)
        fd_c = (loss(a64[b], c64[b] + eps, t64[b]) - loss(a64[b], c64[b] - eps, t64[b])) / (2 * eps)
        scale = max(1.0, np.abs(fd_a).max(), np.abs(fd_t).max(), abs(fd_c))
        assert np.abs(ga[b] - fd_a).max() < 2e-5 * scale, (b, ga[b], fd_a)
        assert np.abs(gt[b] - fd_t).max() < 2e-5 * scale, (b, gt[b], fd_t)
        assert abs(gc[b] - fd_c) < 2e-5 * scale, (b, gc[b], fd_c)
        checked += 1
    assert checked > 0.8 * B
    if penalty == 5000.0:
        assert kinds["inactive"] > 20 and kinds["tight"] > 20
    if penalty <= 2.0:
        assert kinds["relaxed"] > 20


def test_adjoint_structure(coracle):
    """closed-form facts: no gradient flows when the constraint is inactive except through unclipped u_ref; a tight
    constraint makes du orthogonal to a on the free set (a_F . dL/du_ref_F = 0 when r carries no gradient); clipped
    coordinates pass nothing to u_ref"""
    a, c, t, lo, hi, p = cases(7, 2000, 3, 5000.0)
    gu = np.random.default_rng(1).normal(0, 1, a.shape).astype(np.float32)
    ga, gc, gt, margin = coracle.qp_backward(a, c, t, lo, hi, p, gu, None, want_margin=True)
    u, r = coracle.qp(a, c, t, lo, hi, p)
    g0 = c + np.sum(a * np.clip(t, lo, hi), 1)
    inactive = (g0 < -1e-3) & (margin > 1e-3)
    assert inactive.sum() > 100
    assert np.all(ga[inactive] == 0) and np.all(gc[inactive] == 0)
    free = (t > lo) & (t < hi)
    assert np.array_equal(gt[inactive], np.where(free[inactive], gu[inactive], 0))
    tight = (g0 > 1e-3) & (r == 0) & (margin > 1e-3)
    assert tight.sum() > 100
    on_bound = (u <= lo) | (u >= hi)
    assert np.all(gt[tight][on_bound[tight]] == 0)
    assert np.abs(np.sum(a[tight] * gt[tight], 1)).max() < 1e-4


def test_autograd_plumbing_on_fakes(monkeypatch, coracle):
    """ops.qp_diff is a torch.autograd.Function over (qp, qp_backward); CLFController.solve_CLF_QP(requires_grad=True)
    and the neural controller's differentiable path deliver gradients to V / Lf_V / Lg_V / u_ref and to the network"""
    fake_ops.install(monkeypatch)
    from cbfrrt import ops
    a, c, t, lo, hi, p = cases(4, 64, 5, 5000.0)
    A, Cc, T = (torch.tensor(x, requires_grad=True) for x in (a, c, t))
    u, r = ops.qp_diff(A, Cc, T, torch.tensor(lo), torch.tensor(hi), p)
    w = torch.linspace(-1, 1, 4)
    ((u * w).sum() + 0.3 * r.sum()).backward()
    gu = np.tile(w.numpy(), (64, 1)).astype(np.float32)
    ga, gc, gt = coracle.qp_backward(a, c, t, lo, hi, p, gu, np.full(64, 0.3, np.float32))
    assert np.array_equal(A.grad.numpy(), ga) and np.array_equal(Cc.grad.numpy(), gc) and np.array_equal(T.grad.numpy(), gt)

    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    env = ArmEnv(["magician"], config_file="", include_floor=True, device="cpu")
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=env.robot_list[0])
    torch.manual_seed(1)
    ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0,
                                     safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48)
    with torch.no_grad():                              # make the barrier constraint bite on part of the batch
        ctrl.h_nn.output_linear.bias += 30.0
    x = dyn.sample_safe(32)
    (u0, r0), _ = ctrl.solve_CLF_QP(x)
    (u1, r1), _ = ctrl.solve_CLF_QP(x, requires_grad=True)
    assert u1.requires_grad and torch.allclose(u1, u0, atol=2e-5) and torch.allclose(r1, r0, atol=2e-4)
    assert (u0 - dyn.u_nominal(x)).abs().max() > 1e-3  # the filter did change some controls
    (u1.pow(2).sum() + r1.sum()).backward()
    grads = [p.grad for p in ctrl.h_nn.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads) and sum(float(g.abs().sum()) for g in grads) > 0
    # the gradient w.r.t. the output bias is dL/dc * lambda summed over the batch: check it against the oracle adjoint
    h, J, _ = ctrl.h_with_jacobian(x)
    up, lo_ = dyn.control_limits
    ga, gc, gt = coracle.qp_backward(J[:, 0].numpy(), 0.1 * h[:, 0].numpy(), dyn.u_nominal(x).numpy(), lo_.numpy(),
                                     up.numpy(), 5000.0, 2 * u0.numpy(), np.ones(32, np.float32))
    assert abs(float(ctrl.h_nn.output_linear.bias.grad) - 0.1 * gc.sum()) < 1e-3 * max(1.0, abs(0.1 * gc.sum()))


def test_device_code_built_for_the_host_matches_the_oracle(coracle, tmp_path):
    """csrc/mlp_qp.cuh qp_backward<N, double> is __host__ __device__: the very code the kernel runs (incl. the
    float-guided double root find) is compiled for the CPU and compared with the oracle adjoint."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "qp_backward_host")
    subprocess.check_call([nvcc, "-O2", "-w", "-o", exe, os.path.join(root, "tools", "experiments", "qp_backward_host.cu")])
    for n, p in [(7, 5000.0), (4, 0.5), (8, 1e9), (1, 2.0)]:
        a, c, t, lo, hi, _ = cases(n, 20000, 40 + n, p)
        rng = np.random.default_rng(7)
        gu, gr = rng.normal(size=a.shape).astype(np.float32), rng.normal(size=len(c)).astype(np.float32)
        with open(tmp_path / "in", "wb") as fh:
            for x in (a, c, t, gu, gr, lo, hi):
                fh.write(np.ascontiguousarray(x, np.float32).tobytes())
        subprocess.check_call([exe, str(n), str(len(c)), repr(p), str(tmp_path / "in"), str(tmp_path / "out")])
        out = np.fromfile(tmp_path / "out", np.float32)
        B = len(c)
        ga, gc, gt = out[:B * n].reshape(B, n), out[B * n:B * n + B], out[B * n + B:].reshape(B, n)
        ga_o, gc_o, gt_o, margin = coracle.qp_backward(a, c, t, lo, hi, p, gu, gr, want_margin=True)
        ok = margin > 1e-6
        scale = np.maximum(1.0, np.maximum(np.abs(ga_o).max(1), np.maximum(np.abs(gt_o).max(1), np.abs(gc_o))))
        err = np.maximum(np.abs(ga - ga_o).max(1), np.maximum(np.abs(gt - gt_o).max(1), np.abs(gc - gc_o))) / scale
        assert ok.mean() > 0.999 and err[ok].max() < 1e-6, (n, p, err[ok].max())


# ------------------------------------------------------------------------------------------ S > 1 constraints
def kkt_solve64(a, c, t, lo, hi, p, F, clipped_at, W, P):
    """float64 solve of the KKT system on a GIVEN active set (free coordinates F, clipped ones at ``clipped_at``,
    tight constraints W, relaxed constraints P): unknowns (u_F, mu_W).  Independent of the adjoint formulas."""
    n = len(t)
    u = clipped_at.copy()
    nF, nW = int(F.sum()), len(W)
    K = np.zeros((nF + nW, nF + nW))
    rhs = np.zeros(nF + nW)
    K[:nF, :nF] = np.eye(nF)
    AW = a[W][:, F] if nW else np.zeros((0, nF))
    K[:nF, nF:] = 0.5 * AW.T
    K[nF:, :nF] = AW
    rhs[:nF] = t[F] - 0.5 * p * a[P][:, F].sum(0) if len(P) else t[F]
    if nW:
        rhs[nF:] = -c[W] - a[W][:, ~F] @ clipped_at[~F]
    sol = np.linalg.solve(K, rhs)
    u[F] = sol[:nF]
    r = np.zeros(len(c))
    if len(P):
        r[P] = c[P] + a[P] @ u
    return u, r, sol[nF:]


@pytest.mark.parametrize("n,S,penalty", [(7, 2, 50.0), (7, 3, 5000.0), (4, 2, 0.5), (7, 4, 2.0), (3, 2, 50.0)])
def test_multi_adjoint_equals_finite_differences(coracle, n, S, penalty):
    rng = np.random.default_rng(31 * n + S)
    B = 300
    f = lambda x: np.asarray(x, np.float32)
    lo, hi = f(-rng.uniform(0.5, 2.0, n)), f(rng.uniform(0.5, 2.0, n))
    a, c, t = f(rng.normal(size=(B, S, n))), f(rng.normal(scale=0.8, size=(B, S))), f(rng.uniform(-2.0, 2.0, size=(B, n)))
    gu, gr = f(rng.normal(size=(B, n))), f(rng.normal(size=(B, S)))
    ga, gc, gt, margin = coracle.qp_multi_backward(a, c, t, lo, hi, penalty, gu, gr, want_margin=True)
    uo, ro = coracle.qp_multi(a, c, t, lo, hi, penalty, max_iter=5000, tol=1e-13)
    lo64, hi64 = lo.astype(np.float64), hi.astype(np.float64)
    eps, checked, seen = 1e-6, 0, {"tight": 0, "relaxed": 0, "two_tight": 0}
    for b in range(B):
        if margin[b] < 2e-3:
            continue
        a64, c64, t64 = a[b].astype(np.float64), c[b].astype(np.float64), t[b].astype(np.float64)
        u64 = uo[b].astype(np.float64)
        F = (u64 > lo64 + 1e-6) & (u64 < hi64 - 1e-6)
        clipped_at = np.where(u64 <= lo64 + 1e-6, lo64, hi64)
        g = c64 + a64 @ u64
        P = [i for i in range(S) if ro[b, i] > 1e-5]
        W = [i for i in range(S) if ro[b, i] <= 1e-5 and abs(g[i]) < 1e-5]
        if len(W) > F.sum():
            continue
        u0, r0, mu0 = kkt_solve64(a64, c64, t64, lo64, hi64, min(penalty, 1e6), F, clipped_at, W, P)
        assert np.abs(u0 - u64).max() < 1e-5 and np.all(mu0 > 0) and np.all(mu0 < min(penalty, 1e6))
        seen["tight"] += len(W) >= 1
        seen["two_tight"] += len(W) >= 2
        seen["relaxed"] += len(P) >= 1

        def loss(av, cv, tv):
            u, r, _ = kkt_solve64(av, cv, tv, lo64, hi64, min(penalty, 1e6), F, clipped_at, W, P)
            return float(gu[b].astype(np.float64) @ u + gr[b].astype(np.float64) @ r)
        fd_a, fd_c, fd_t = np.zeros((S, n)), np.zeros(S), np.zeros(n)
        for i in range(S):
            d = np.zeros(S); d[i] = eps
            fd_c[i] = (loss(a64, c64 + d, t64) - loss(a64, c64 - d, t64)) / (2 * eps)
            for k in range(n):
                d2 = np.zeros((S, n)); d2[i, k] = eps
                fd_a[i, k] = (loss(a64 + d2, c64, t64) - loss(a64 - d2, c64, t64)) / (2 * eps)
        for k in range(n):
            d = np.zeros(n); d[k] = eps
            fd_t[k] = (loss(a64, c64, t64 + d) - loss(a64, c64, t64 - d)) / (2 * eps)
        scale = max(1.0, np.abs(fd_a).max(), np.abs(fd_c).max(), np.abs(fd_t).max())
        assert np.abs(ga[b] - fd_a).max() < 3e-5 * scale, (b, ga[b], fd_a)
        assert np.abs(gc[b] - fd_c).max() < 3e-5 * scale, (b, gc[b], fd_c)
        assert np.abs(gt[b] - fd_t).max() < 3e-5 * scale, (b, gt[b], fd_t)
        checked += 1
    assert checked > 0.6 * B and seen["tight"] > 20
    if penalty <= 2.0:
        assert seen["relaxed"] > 20
    if penalty >= 50.0 and n >= 4:
        assert seen["two_tight"] > 5


def test_multi_adjoint_reduces_to_single(coracle):
    a, c, t, lo, hi, p = cases(7, 500, 8, 50.0)
    rng = np.random.default_rng(2)
    gu, gr = rng.normal(size=a.shape).astype(np.float32), rng.normal(size=len(c)).astype(np.float32)
    ga1, gc1, gt1, m1 = coracle.qp_backward(a, c, t, lo, hi, p, gu, gr, want_margin=True)
    gaS, gcS, gtS, mS = coracle.qp_multi_backward(a[:, None], c[:, None], t, lo, hi, p, gu, gr[:, None], want_margin=True)
    ok = (m1 > 1e-4) & (mS > 1e-4)
    assert ok.mean() > 0.99
    assert np.abs(gaS[ok, 0] - ga1[ok]).max() < 1e-4 and np.abs(gcS[ok, 0] - gc1[ok]).max() < 1e-4
    assert np.abs(gtS[ok] - gt1[ok]).max() < 1e-4


def test_multi_device_code_built_for_the_host_matches_the_oracle(coracle, tmp_path):
    """qp_solve_multi + qp_multi_adjoint (csrc/mlp_qp.cuh, __host__ __device__) compiled for the CPU vs the oracle adjoint
    (which classifies the active set from the multipliers; the device code reads it off the solution)"""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "qp_multi_backward_host")
    subprocess.check_call([nvcc, "-O2", "-w", "-o", exe, os.path.join(root, "tools", "experiments", "qp_multi_backward_host.cu")])
    f = lambda x: np.ascontiguousarray(x, np.float32)
    for n, S, p in [(7, 2, 50.0), (7, 4, 0.05), (4, 3, 5000.0), (8, 3, 1e9)]:
        rng = np.random.default_rng(300 + n * S)
        B = 5000
        lo, hi = f(-rng.uniform(0.5, 2.0, n)), f(rng.uniform(0.5, 2.0, n))
        a, c, ur = f(rng.normal(size=(B, S, n))), f(rng.normal(scale=0.8, size=(B, S))), f(rng.uniform(-2.0, 2.0, size=(B, n)))
        gu, gr = f(rng.normal(size=(B, n))), f(rng.normal(size=(B, S)))
        with open(tmp_path / "in", "wb") as fh:
            for x in (a, c, ur, gu, gr, lo, hi):
                fh.write(x.tobytes())
        subprocess.check_call([exe, str(n), str(S), str(B), repr(p), str(tmp_path / "in"), str(tmp_path / "out")])
        out = np.fromfile(tmp_path / "out", np.float32)
        ga, gc, gt = (out[:B * S * n].reshape(B, S, n), out[B * S * n:B * S * n + B * S].reshape(B, S),
                      out[B * S * n + B * S:].reshape(B, n))
        ga_o, gc_o, gt_o, margin = coracle.qp_multi_backward(a, c, ur, lo, hi, p, gu, gr, want_margin=True)
        scale = np.maximum(1.0, np.maximum(np.abs(ga_o).max((1, 2)), np.maximum(np.abs(gt_o).max(1), np.abs(gc_o).max(1))))
        err = np.maximum(np.abs(ga - ga_o).max((1, 2)), np.maximum(np.abs(gt - gt_o).max(1), np.abs(gc - gc_o).max(1))) / scale
        ok = margin > 1e-5
        assert ok.mean() > 0.999 and err[ok].max() < 1e-5, (n, S, p, err[ok].max())


def test_alternative_backends_keep_the_reference_signatures(monkeypatch, coracle):
    """_solve_CLF_QP_cvxpylayers / _gurobi / _qpsolver (clf_controller.py:226-459): same arguments, same (u, r) shapes,
    the documented semantics of each, all on the exact solver"""
    fake_ops.install(monkeypatch)
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    env = ArmEnv(["magician"], config_file="", include_floor=True, device="cpu")
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=env.robot_list[0])
    torch.manual_seed(1)
    ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0,
                                     safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48)
    with torch.no_grad():
        ctrl.h_nn.output_linear.bias += 30.0
    x = dyn.sample_safe(24)
    V, Lf, Lg, _ = ctrl.V_with_lie_derivatives(x)
    u_ref = dyn.u_nominal(x)
    (u0, r0), _ = ctrl.solve_CLF_QP(x)
    u1, r1 = ctrl._solve_CLF_QP_cvxpylayers(x, u_ref, V, Lf, Lg, 5000.0)
    assert u1.shape == (24, 4) and r1.shape == (24, 1) and torch.allclose(u1, u0, atol=2e-5) and torch.allclose(r1, r0, atol=2e-4)
    u2, r2 = ctrl._solve_CLF_QP_gurobi(x, u_ref, V, Lf, Lg, 5000.0)
    assert torch.allclose(u2, u1, atol=1e-6) and torch.allclose(r2, r1, atol=1e-6)
    xb = x.clone(); xb[3, 0] = float("nan")
    Lgb = Lg.clone(); Lgb[5, 0, 1] = float("inf")
    u3, r3 = ctrl._solve_CLF_QP_gurobi(xb, u_ref, V, Lf, Lgb, 5000.0)
    assert (u3[3] == 0).all() and (u3[5] == 0).all() and r3[3] == 0 and torch.allclose(u3[6:], u1[6:], atol=1e-6)
    # penalty form of the OSQP path: u = clip(u_ref - p Lg / 2), r = 0
    up, lo = dyn.control_limits
    u4, r4 = ctrl._solve_CLF_QP_qpsolver(x, u_ref, V, Lf, Lg, 0.3)
    exp = torch.max(torch.min(u_ref - 0.15 * Lg[:, 0, :], up), lo)
    assert torch.allclose(u4, exp, atol=1e-6) and r4.shape == V.shape and (r4 == 0).all()
    # hard-constrained form: feasible batch -> the constrained optimum (== relaxed QP with a huge penalty, r = 0)
    u5, r5 = ctrl._solve_CLF_QP_qpsolver(x, u_ref, V, Lf, Lg, 0.3, use_constraint=True)
    ub, rb = ctrl._solve_CLF_QP_exact(x, u_ref, V, Lf, Lg, 1e6)
    if float(rb.max()) <= 1e-6:
        assert torch.allclose(u5, ub, atol=1e-6)
        g = (Lf[:, 0, 0] + (Lg[:, 0, :] * u5).sum(1) + 0.1 * V.reshape(-1))
        assert float(g.max()) <= 1e-4
    else:                         # an infeasible row: whole batch falls back to the penalty form (:449-452)
        assert torch.allclose(u5, exp, atol=1e-6)
    # an infeasible hard constraint forces the fallback
    u6, _ = ctrl._solve_CLF_QP_qpsolver(x, u_ref, V + 1e4, Lf, Lg, 0.3, use_constraint=True)
    assert torch.allclose(u6, exp, atol=1e-6)
