"""CPU suite: pins the oracles against every known answer the reference's text yields and against
each other (float32 contract C oracle vs independent float64 derivation)."""
import numpy as np
import pytest

from conftest import rel_err


def test_fk_known_answers_q0_and_zero(models, coracle):
    """Panda link origins at q = 0 and link8 at q0, computed by hand from panda.urdf."""
    from oracle import oracle_f64 as o
    m = models["panda"]
    expect0 = {0: (0, 0, 0.333), 1: (0, 0, 0.333), 2: (0, 0, 0.649), 3: (0.0825, 0, 0.649), 4: (0, 0, 1.033),
               5: (0, 0, 1.033), 6: (0.088, 0, 1.033), 7: (0.088, 0, 0.926)}
    fr = o.fk(m, np.zeros(7))
    pos, rot = coracle.fk("panda", np.zeros((1, 7), np.float32))
    for l, e in expect0.items():
        assert np.allclose(fr[l][:3, 3], e, atol=1e-9)
        assert np.allclose(pos[0, l], e, atol=1e-6)
    fr = o.fk(m, np.array(m["q0"]))
    assert np.allclose(fr[7][:3, 3], (0.288556, -0.257364, 0.361447), atol=1e-6)
    pos, _ = coracle.fk("panda", np.array([m["q0"]], np.float32))
    assert np.allclose(pos[0, 7], (0.288556, -0.257364, 0.361447), atol=2e-6)


def test_urdf_constants(models):
    p, g = models["panda"], models["magician"]
    assert p["n"] == 7 and p["n_links"] == 12 and p["body_joints"] == list(range(7)) and p["ee_joints"] == [9, 10]
    assert np.allclose(p["q_lower"], [-2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671])
    assert np.allclose(p["q_upper"], [2.9671, 1.8326, 2.9671, 0.0, 2.9671, 3.8223, 2.9671])
    assert np.allclose(p["v_max"], [2.175] * 4 + [2.61] * 3)
    assert g["n"] == 4 and g["n_links"] == 4 and np.allclose(g["v_max"], [3.15] * 4)
    assert np.allclose(g["q_upper"], [3.14159265, 1.570796325, 1.570796325, 3.14159265])
    # magician joint origins carry globalScaling = 2 (environment/magician.py:20)
    assert np.allclose(g["links"][0]["t64"], [0, 0, 0.048]) and np.allclose(g["links"][1]["t64"], [-0.0235, 0, 0.228])
    assert np.allclose(g["q0"], [0, 0.707, 0.707, 0])


def test_lqr_gain_known_values(golden):
    from cbfrrt.neural_cbf.systems.utils import scalar_dare_gain, lqr
    import scipy.linalg
    assert abs(scalar_dare_gain(1 / 30) - 0.9834722125784985) < 1e-13
    assert abs(scalar_dare_gain(0.01) - 0.9950124999218958) < 1e-13
    for n, b in ((7, 1 / 30), (4, 0.01)):
        A, B = np.eye(n), b * np.eye(n)
        X = scipy.linalg.solve_discrete_are(A, B, np.eye(n), np.eye(n))
        K_ref = np.linalg.inv(B.T @ X @ B + np.eye(n)) @ (B.T @ X @ A)  # reference utils.py:40-43
        assert np.allclose(lqr(A, B, np.eye(n), np.eye(n)), K_ref, atol=1e-9)
    # K produced by the reference's own compute_linearized_controller (golden, scipy DARE)
    for name in ("panda", "magician"):
        K = golden[name]["K"]
        assert np.allclose(K, scalar_dare_gain(1 / 30) * np.eye(K.shape[0]), atol=1e-8)


def test_contract_sincos_accuracy(coracle):
    x = np.linspace(-7.0, 7.0, 200001).astype(np.float32)
    s, c = coracle.sincos(x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 2.5e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 2.5e-7


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_c_oracle_matches_f64_oracle(models, coracle, robot):
    """float32-contract oracle vs the independent float64 derivation: identical booleans/indices away from
    threshold ties, values within 1e-5 relative."""
    from oracle import oracle_f64 as o
    m = models[robot]
    rng = np.random.default_rng(11)
    q = rng.uniform(m["q_lower"], m["q_upper"], size=(300, m["n"])).astype(np.float32)
    boxes = np.array([[0.3, 0.17, 0.3, 0.05, 0.05, 0.1], [-0.4, 0.23, 0.4, 0.05, 0.05, 0.1],
                      [0.0, -0.23, 0.6, 0.05, 0.05, 0.1]])
    sph = np.array([[0.3, -0.3, 0.5, 0.1]])
    out = coracle.collide(robot, q, boxes, sph, True, 0.02, 0.0)
    pos, rot = coracle.fk(robot, q)
    n = m["n"]
    for b in range(q.shape[0]):
        fr = o.fk(m, q[b].astype(np.float64))
        for l in range(m["n_links"]):
            assert rel_err(pos[b, l], fr[l][:3, 3]) < 1e-5 and rel_err(rot[b, l], fr[l][:3, :3]) < 1e-5
        r = o.evaluate(m, q[b].astype(np.float64), boxes, sph, True, 0.02)
        near_tie = min(abs(r["soft_min"] - 0.02), abs(r["hard_min"]), abs(r["self_min"] - 0.01)) < 1e-5
        if not near_tie:
            assert bool(out["safe"][b]) == r["safe"] and bool(out["unsafe"][b]) == r["unsafe"]
        assert abs(out["datax"][b, n] - r["o"]) < 1e-5
        assert abs(out["mins"][b, 1] - r["soft_min"]) < 1e-5 and abs(out["mins"][b, 2] - r["hard_min"]) < 1e-5
        assert abs(out["mins"][b, 0] - r["self_min"]) < 1e-5
        if out["arg_obs"][b] == r["arg_obs"] and out["arg_link"][b] == m["spheres"][r["arg_sphere"]]["link"]:
            assert rel_err(out["datax"][b, n + 1:], r["do_dq"]) < 1e-5


def test_gradient_is_derivative_of_distance(models, coracle):
    """do/dq returned with the observation is the derivative of o along q (central differences)."""
    m = models["panda"]
    rng = np.random.default_rng(5)
    boxes = np.array([[0.35, 0.1, 0.4, 0.1, 0.1, 0.1], [-0.3, -0.3, 0.6, 0.08, 0.1, 0.15]])
    q = rng.uniform(m["q_lower"], m["q_upper"], size=(200, 7)).astype(np.float32)
    base = coracle.collide("panda", q, boxes, None, True)
    eps, ok, tot = 1e-3, 0, 0
    for j in range(7):
        qp, qm = q.copy(), q.copy()
        qp[:, j] += eps
        qm[:, j] -= eps
        op, om = coracle.collide("panda", qp, boxes, None, True), coracle.collide("panda", qm, boxes, None, True)
        same = (op["arg_obs"] == base["arg_obs"]) & (om["arg_obs"] == base["arg_obs"]) & \
               (op["arg_link"] == base["arg_link"]) & (om["arg_link"] == base["arg_link"])
        fd = (op["datax"][:, 7] - om["datax"][:, 7]) / (2 * eps)
        err = np.abs(fd - base["datax"][:, 8 + j])[same]
        ok += (err < 5e-3).sum()
        tot += same.sum()
    assert tot > 500 and ok / tot > 0.97


def test_edge_cases_empty_scene_and_no_floor(coracle):
    q = np.zeros((3, 7), np.float32)
    out = coracle.collide("panda", q, None, None, False)
    assert np.isinf(out["datax"][:, 7]).all() and (out["arg_obs"] == -1).all() and (out["unsafe"] == 0).all()
    out = coracle.collide("panda", q, None, None, True)
    # link 0 rests 0.141 m above the floor: the kinematic constant wins at q = 0 (upright arm)
    assert np.allclose(out["datax"][:, 7], 0.14099599) and (out["arg_obs"] == 0).all() and (out["arg_link"] == 0).all()
    assert (out["datax"][:, 8:] == 0).all()


def test_sampler_is_shard_invariant(coracle):
    a = coracle.sample_states("panda", 7, 0, 1000)
    b = np.concatenate([coracle.sample_states("panda", 7, 0, 400), coracle.sample_states("panda", 7, 400, 600)])
    assert np.array_equal(a, b)
    lo, hi = np.array([-2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671]), \
        np.array([2.9671, 1.8326, 2.9671, 0.0, 2.9671, 3.8223, 2.9671])
    assert (a >= lo - 1e-6).all() and (a <= hi + 1e-6).all() and a.std(0).min() > 0.4


def test_qp_multi_oracle_is_optimal(coracle):
    """oracle_qp_multi (dual coordinate ascent, fp64) on the QP of clf_controller.py:86-139 with S scenarios: KKT
    certificate for every row, and agreement with scipy SLSQP on the primal wherever SLSQP converges"""
    from scipy.optimize import minimize
    from conftest import qp_kkt_residual
    rng = np.random.default_rng(7)
    n_slsqp = 0
    for n, S, p in ((4, 3, 50.0), (7, 4, 1e3), (4, 2, 0.05), (7, 2, 1e9), (4, 4, 1e9)):
        B = 60
        lo, hi = -np.ones(n, np.float32), np.ones(n, np.float32)
        a = rng.normal(size=(B, S, n)).astype(np.float32)
        c = rng.normal(scale=0.8, size=(B, S)).astype(np.float32)
        ur = rng.uniform(-1.5, 1.5, size=(B, n)).astype(np.float32)      # also outside the box
        u, r, it = coracle.qp_multi(a, c, ur, lo, hi, p, want_iters=True)
        assert it.max() <= 12, it.max()          # active-set identification: a handful of iterations
        pc = min(p, 1e6)
        for b in range(B):
            g = c[b].astype(np.float64) + a[b].astype(np.float64) @ u[b]
            np.testing.assert_allclose(r[b], np.maximum(g, 0), atol=1e-5 * (1 + np.abs(c[b]).max()))
            assert np.all(u[b] >= lo) and np.all(u[b] <= hi)
            assert qp_kkt_residual(a[b], c[b], ur[b], lo, hi, pc, u[b]) <= 1e-4 * (1 + pc * 1e-3), (n, S, p, b)
            if p > 100:
                continue
            cons = [{"type": "ineq", "fun": (lambda z, i=i: -(c[b, i] + a[b, i] @ z[:n] - z[n + i]))} for i in range(S)]
            u0 = np.clip(ur[b], lo, hi)
            res = minimize(lambda z: np.sum((z[:n] - ur[b]) ** 2) + p * np.sum(z[n:]),
                           np.concatenate([u0, np.maximum(0, c[b] + a[b] @ u0) + 0.1]), constraints=cons,
                           bounds=[(-1, 1)] * n + [(0, None)] * S, method="SLSQP", options={"ftol": 1e-12, "maxiter": 500})
            if res.success:
                n_slsqp += 1
                np.testing.assert_allclose(u[b], res.x[:n], atol=3e-4)
    assert n_slsqp > 60


def test_qp_single_reference_outside_box(coracle):
    """u_ref outside the control box: a coordinate leaves one bound and can reach the other (2n breakpoints)"""
    from conftest import qp_kkt_residual
    rng = np.random.default_rng(11)
    n, B, p = 4, 200, 1e3
    lo, hi = -np.ones(n, np.float32), np.ones(n, np.float32)
    a = rng.normal(size=(B, n)).astype(np.float32)
    c = rng.normal(scale=1.0, size=B).astype(np.float32)
    ur = rng.uniform(-3, 3, size=(B, n)).astype(np.float32)
    u, r = coracle.qp(a, c, ur, lo, hi, p)
    um, rm = coracle.qp_multi(a[:, None, :], c[:, None], ur, lo, hi, p)
    np.testing.assert_array_equal(u, um)
    np.testing.assert_allclose(r, rm[:, 0], atol=1e-6)
    for b in range(B):
        assert qp_kkt_residual(a[b:b + 1], c[b:b + 1], ur[b], lo, hi, p, u[b]) <= 1e-3


def test_knn_oracle_matches_numpy(coracle):
    """oracle_knn against the reference's numpy expressions (batch_tsa.py:135-145): argmin per row, argsort[:k]"""
    rng = np.random.default_rng(3)
    for n in (4, 7):
        tree = rng.uniform(-2, 2, size=(700, n)).astype(np.float32)
        samples = rng.uniform(-2, 2, size=(40, n)).astype(np.float32)
        tree[100] = tree[7]                       # duplicate node: ties go to the lowest index
        samples[0] = tree[7]
        d = np.linalg.norm(samples[:, None, :].astype(np.float64) - tree[None, :, :], axis=-1)
        idx, dist = coracle.knn(samples, tree, 1)
        assert np.array_equal(idx[:, 0], np.argmin(d, axis=1)) and idx[0, 0] == 7
        np.testing.assert_allclose(dist[:, 0], d.min(1), rtol=2e-6, atol=1e-7)
        k = 50
        idx, dist = coracle.knn(samples[1:2], tree, k)
        order = np.argsort(d[1], kind="stable")[:k]
        assert np.array_equal(idx[0], order)
        np.testing.assert_allclose(dist[0], np.sort(d[1])[:k], rtol=2e-6)
        idx, dist = coracle.knn(samples[:3], tree[:5], 8)          # k larger than the tree
        assert (idx[:, 5:] == -1).all() and np.isinf(dist[:, 5:]).all() and (np.sort(idx[:, :5], 1) == np.arange(5)).all()
        within = coracle.radius_mask(samples, tree, 1.3)
        assert np.mean(within.astype(bool) == (d <= 1.3)) > 0.9999   # fp32 vs fp64 at the boundary


def test_qp_multi_degenerate_constraints(coracle):
    """duplicate / zero / (anti-)parallel constraint rows (singular Newton matrices, flat dual ridges): finite, feasible,
    KKT-optimal controls within a handful of iterations"""
    from conftest import qp_kkt_residual, degenerate_qp_batch
    a, c, ur = degenerate_qp_batch()
    lo, hi = -np.ones(4, np.float32), np.ones(4, np.float32)
    for p in (50.0, 1e9):
        u, r, it = coracle.qp_multi(a, c, ur, lo, hi, p, max_iter=300, want_iters=True)
        assert np.isfinite(u).all() and np.isfinite(r).all() and (np.abs(u) <= 1).all() and (r >= 0).all()
        assert it.max() <= 12, it.max()
        np.testing.assert_allclose(r[500:1000, 2], np.maximum(c[500:1000, 2], 0), atol=1e-5)
        rows = list(range(0, a.shape[0], 9)) + list(np.where(it > 12)[0])
        assert max(qp_kkt_residual(a[b], c[b], ur[b], lo, hi, min(p, 1e6), u[b]) for b in rows) < 1e-4 * max(1, min(p, 1e6) * 1e-3)


def test_qp_single_edge_cases(coracle):
    """zero / partly zero / tiny / huge gradients, u_ref on the box, asymmetric limits, penalties 0 .. 1e9 (clamped):
    finite, inside the box, r >= 0, KKT-optimal"""
    from conftest import qp_kkt_residual, edge_case_qp_batch
    a, c, ur, lo, hi = edge_case_qp_batch()
    for p in (0.0, 0.05, 5000.0, 1e9):
        u, r = coracle.qp(a, c, ur, lo, hi, p)
        assert np.isfinite(u).all() and np.isfinite(r).all() and (u >= lo).all() and (u <= hi).all() and (r >= 0).all()
        if p > 0:
            np.testing.assert_allclose(r[:300], np.maximum(c[:300], 0), atol=1e-6)      # a = 0: only relaxation helps
        assert max(qp_kkt_residual(a[b:b + 1], c[b:b + 1], ur[b], lo, hi, min(p, 1e6), u[b]) for b in range(0, len(c), 3)) < 1e-5


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_scene_permutation_and_monotonicity_properties(coracle, robot):
    """size-independent properties of the distance stage: the minimum and the masks do not depend on the obstacle
    order (the arg-min follows the permutation), and adding an obstacle can only lower the minimum / clear `safe`"""
    from cbfrrt import workloads as W
    q = W.sample_states(robot, 3000, 23)
    boxes, spheres = W.random_boxes(9, 12, 0.03, 0.12), W.random_spheres(4, 13, 0.03, 0.1)
    n = q.shape[1]
    base = coracle.collide(robot, q, boxes, spheres, True, 0.02, 0.0)
    perm = np.random.default_rng(1).permutation(len(boxes))
    pr = coracle.collide(robot, q, boxes[perm], spheres, True, 0.02, 0.0)
    assert np.array_equal(base["safe"], pr["safe"]) and np.array_equal(base["unsafe"], pr["unsafe"])
    assert np.array_equal(base["datax"][:, n], pr["datax"][:, n])
    is_box = (base["arg_obs"] >= 1) & (base["arg_obs"] <= len(boxes))
    assert np.array_equal(perm[pr["arg_obs"][is_box] - 1], base["arg_obs"][is_box] - 1)
    more = coracle.collide(robot, q, np.concatenate([boxes, W.random_boxes(3, 14, 0.05, 0.15)]), spheres, True, 0.02, 0.0)
    assert np.all(more["datax"][:, n] <= base["datax"][:, n]) and not np.any(more["safe"] & ~base["safe"])
    assert not np.any(base["unsafe"] & ~more["unsafe"])
    assert np.any(more["datax"][:, n] < base["datax"][:, n])


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_geometry_contract_fixture(coracle, robot):
    """tests/golden/geometry_contract_*.npz (make_geometry_golden.py): the sphere model, the exemption rules and the
    float32 operation order have not drifted -- bit for bit"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"geometry_contract_{robot}.npz"))
    for floor in (1, 0):
        o = coracle.collide(robot, g["q"], g["boxes"], g["spheres"], bool(floor), 0.02, 0.0)
        for k in ("safe", "unsafe", "datax", "arg_link", "arg_obs"):
            assert np.array_equal(o[k], g[f"{k}_floor{floor}"]), (k, floor)
    pos, rot = coracle.fk(robot, g["q"][:50])
    assert np.array_equal(pos, g["fk_pos"]) and np.array_equal(rot, g["fk_rot"])


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_oracle_matches_reference_systems_golden(coracle, robot):
    """the C oracle against the outputs of the reference's own safe_mask / unsafe_mask /
    complete_sample_with_observations loops (tests/golden/make_ref_systems_golden.py)"""
    from conftest import check_against_reference_systems_golden
    check_against_reference_systems_golden(robot, lambda q, boxes, floor: coracle.collide(robot, q, boxes, None, floor, 0.02, 0.0))
