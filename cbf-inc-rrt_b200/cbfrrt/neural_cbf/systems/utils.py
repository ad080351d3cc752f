"""LQR / Lyapunov helpers -- mirror of neural_cbf/systems/utils.py:18-70 for the systems on the hot path.

The reference calls scipy.linalg.solve_discrete_are (utils.py:40) for every ``set_goal`` /
``set_intermediate_goals`` (arm_dynamics.py:64-74), i.e. once per steer call.  For the velocity-controlled arms
A = I, B = dt_c * I, Q = R = I, so the DARE decouples into n identical scalar equations with the closed form
below (matches scipy to 1e-15; pinned by tests/golden/ref_controller_*.npz, K[0,0] = 0.98347221...).
"""
from typing import Dict, List

import numpy as np

grav = 9.80665
Scenario = Dict[str, float]
ScenarioList = List[Scenario]


def scalar_dare_gain(b: float) -> float:
    """x = 1 + x - (bx)^2/(1 + b^2 x)  ->  b^2 x^2 - b^2 x - 1 = 0;  K = b x / (1 + b^2 x)."""
    b2 = b * b
    x = (b2 + np.sqrt(b2 * b2 + 4.0 * b2)) / (2.0 * b2)
    return b * x / (1.0 + b2 * x)


def lqr(A: np.ndarray, B: np.ndarray, Q: np.ndarray, R: np.ndarray, return_eigs: bool = False):
    """Discrete-time LQR gain, u = -K x (utils.py:18-49).  Closed form when A = I, B = b I, Q = R = I; general
    matrices fall back to scipy (host-side, set-up time only, exactly what the reference does)."""
    n = A.shape[0]
    if (A.shape == B.shape and np.array_equal(A, np.eye(n)) and np.array_equal(Q, np.eye(n))
            and np.array_equal(R, np.eye(n)) and np.array_equal(B, B[0, 0] * np.eye(n)) and B[0, 0] != 0):
        K = scalar_dare_gain(float(B[0, 0])) * np.eye(n)
    else:
        import scipy.linalg
        X = scipy.linalg.solve_discrete_are(A, B, Q, R)
        K = np.linalg.inv(B.T @ X @ B + R) @ (B.T @ X @ A)
    if not return_eigs:
        return K
    return K, np.linalg.eigvals(A - B @ K)


def continuous_lyap(Acl: np.ndarray, Q: np.ndarray):
    """Acl^T P + P Acl + Q = 0 (utils.py:52-60).  Closed form for Acl = -k I."""
    n = Acl.shape[0]
    if np.array_equal(Acl, Acl[0, 0] * np.eye(n)) and Acl[0, 0] < 0 and np.array_equal(Q, np.eye(n)):
        return np.eye(n) / (-2.0 * Acl[0, 0])
    import scipy.linalg
    return scipy.linalg.solve_continuous_lyapunov(Acl.T, -Q)
