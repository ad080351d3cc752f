"""CPU suite: the C oracle's MLP / Jacobian / Lie-derivative / nominal-control / QP restatement against vectors
produced by the REFERENCE'S OWN PYTHON (tests/golden/make_golden.py) -- the one part of the path that is pinned."""
import numpy as np
import pytest

from conftest import golden_state_dict, rel_err

TOL = 1e-5  # north-star tolerance for barrier values, gradients and QP controls


def _run(coracle, g, name, tag, penalty=None, u_ref=None):
    from cbfrrt.workloads import pack_weights_np
    n = g["datax"].shape[1] // 2
    w = pack_weights_np(golden_state_dict(g), n)
    return coracle.cbf_filter(n, w, 48, 2, g["datax"], g[f"{tag}_goal"], g["K"], g["u_lower"], g["u_upper"],
                              float(g["lam"]), float(g["penalty"]) if penalty is None else penalty, u_ref=u_ref), n


def test_param_counts(golden):
    assert sum(v.size for v in golden_state_dict(golden["panda"]).values()) == 5185
    assert sum(v.size for v in golden_state_dict(golden["magician"]).values()) == 5041


@pytest.mark.parametrize("robot", ["panda", "magician"])
@pytest.mark.parametrize("tag", ["shared", "rows"])
def test_h_jacobian_lie_and_control_match_reference_python(coracle, golden, robot, tag):
    g = golden[robot]
    out, n = _run(coracle, g, robot, tag)
    assert rel_err(out["h"], g[f"{tag}_h"][:, 0], floor=0.1) < TOL
    assert rel_err(out["J"], g[f"{tag}_J"][:, 0, :], floor=0.1) < TOL
    # V_with_lie_derivatives: V = h, Lf = 0, Lg = J (f = 0, g = I)
    assert np.array_equal(g[f"{tag}_V"], g[f"{tag}_h"]) and not g[f"{tag}_Lf"].any()
    assert rel_err(out["J"], g[f"{tag}_Lg"][:, 0, :], floor=0.1) < TOL
    # rows where the filter is inactive return u_ref = clamp(-K(q - goal)): pins u_nominal + goal plumbing
    inactive = np.abs(g[f"{tag}_u"] - g[f"{tag}_u_ref"]).max(1) < 1e-9
    assert inactive.sum() > 10 and (~inactive).sum() > 10
    assert rel_err(out["u"][inactive], g[f"{tag}_u_ref"][inactive]) < TOL
    # active rows: exact KKT solution vs the independent SLSQP solve of the reference's QP parameters
    assert rel_err(out["u"], g[f"{tag}_u"]) < TOL
    assert np.abs(out["r"] - g[f"{tag}_r"][:, 0]).max() < TOL


@pytest.mark.parametrize("robot", ["panda", "magician"])
def test_relaxation_branch_and_penalty_clamp(coracle, golden, robot):
    g = golden[robot]
    out, n = _run(coracle, g, robot, "shared", penalty=0.05)
    assert (g["shared_r_pen0.05"] > 1e-6).sum() > 10
    assert rel_err(out["u"], g["shared_u_pen0.05"]) < TOL
    assert rel_err(out["r"], g["shared_r_pen0.05"][:, 0], floor=0.1) < TOL
    out, n = _run(coracle, g, robot, "shared", penalty=1e9)  # clamped to 1e6 (clf_controller.py:380)
    assert rel_err(out["u"], g["shared_u_pen1e9"]) < TOL


def test_qp_kkt_conditions_random(coracle):
    """Self-checking: stationarity, feasibility and complementarity of the oracle's QP solution."""
    rng = np.random.default_rng(3)
    n, B = 7, 4000
    from cbfrrt.workloads import make_state_dict, pack_weights_np
    sd, _ = make_state_dict(n, seed=4)
    w = pack_weights_np(sd, n)
    lo, hi = -np.full(n, 2.0, np.float32), np.full(n, 2.0, np.float32)
    datax = np.concatenate([rng.uniform(-2, 2, (B, n)), rng.uniform(-0.1, 0.3, (B, 1)), 3 * rng.normal(size=(B, n))], 1)
    K = np.eye(n, dtype=np.float32)
    for pen in (0.5, 50.0, 5000.0):
        out = coracle.cbf_filter(n, w, 48, 2, datax, np.zeros((1, n)), K, lo, hi, 1.0, pen)
        a, u, r, h = out["J"].astype(np.float64), out["u"].astype(np.float64), out["r"].astype(np.float64), out["h"]
        uref = np.clip(-datax[:, :n], lo, hi)
        g = 1.0 * h + (a * u).sum(1) - r
        assert (g < 1e-5).all() and (r > -1e-7).all() and (u >= lo - 1e-6).all() and (u <= hi + 1e-6).all()
        # mu from the free coordinates: 2(u - uref) + mu a = 0
        free = (u > lo + 1e-5) & (u < hi - 1e-5) & (np.abs(a) > 1e-3)
        mu = np.where(free, -2 * (u - uref) / np.where(free, a, 1), np.nan)
        mu_row = np.nanmedian(mu, axis=1)
        has = ~np.isnan(mu_row)
        assert np.nanmax(np.abs(mu - mu_row[:, None])[has]) < 1e-3            # one multiplier per row
        assert (mu_row[has] > -1e-4).all() and (mu_row[has] < pen * (1 + 1e-4) + 1e-4).all()
        active = has & (mu_row > 1e-4)
        assert (np.abs(g[active & (r < 1e-9)]) < 1e-4).all()                    # complementarity
        assert (np.abs(mu_row[has & (r > 1e-6)] - pen) < 1e-3 * pen).all()      # relaxed  <=>  mu = p
        assert active.sum() > 100
