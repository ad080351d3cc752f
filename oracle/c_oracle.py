"""TEST INFRASTRUCTURE -- ctypes/numpy front-end of oracle/cbfrrt_oracle.c (see that file's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "libcbfrrt_oracle.so")
ROBOT_ID = {"panda": 0, "magician": 1}


def build(force=False):
    src = os.path.join(_DIR, "cbfrrt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_DIR, "robot_tables.h"))):
        subprocess.check_call(["make", "-C", _DIR, "-B"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def set_threads(t):
    lib().oracle_set_threads(int(t))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def pack_scene(boxes, spheres=None):
    """boxes (Ob,6) [c, h] -> (Ob,8) device layout; spheres (Os,4)."""
    boxes = np.zeros((0, 6), np.float32) if boxes is None else _f(boxes).reshape(-1, 6)
    b8 = np.zeros((boxes.shape[0], 8), np.float32)
    b8[:, :6] = boxes
    sp = np.zeros((0, 4), np.float32) if spheres is None else _f(spheres).reshape(-1, 4)
    return b8, sp


def robot_info(robot):
    n, l, s = C.c_int(), C.c_int(), C.c_int()
    lib().oracle_robot_info(ROBOT_ID[robot], C.byref(n), C.byref(l), C.byref(s))
    return n.value, l.value, s.value


def sincos(x):
    x = _f(x)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().oracle_sincos(_p(x), C.c_long(x.size), _p(s), _p(c))
    return s, c


def fk(robot, q):
    q = _f(q)
    n, L, _ = robot_info(robot)
    B = q.shape[0]
    pos = np.empty((B, L, 3), np.float32)
    rot = np.empty((B, L, 3, 3), np.float32)
    lib().oracle_fk(ROBOT_ID[robot], _p(q), C.c_long(q.shape[1]), C.c_long(B), _p(pos), _p(rot))
    return pos, rot


def collide(robot, q, boxes, spheres=None, floor=True, thr_soft=0.02, thr_hard=0.0):
    q = _f(q)
    n, L, _ = robot_info(robot)
    B = q.shape[0]
    b8, sp = pack_scene(boxes, spheres)
    out = dict(safe=np.empty(B, np.uint8), unsafe=np.empty(B, np.uint8), datax=np.empty((B, 2 * n + 1), np.float32),
               arg_link=np.empty(B, np.int32), arg_obs=np.empty(B, np.int32), mins=np.empty((B, 3), np.float32))
    lib().oracle_collide(ROBOT_ID[robot], b8.shape[0], _p(b8), sp.shape[0], _p(sp), int(bool(floor)), _p(q),
                         C.c_long(q.shape[1]), C.c_long(B), C.c_float(thr_soft), C.c_float(thr_hard), _p(out["safe"]),
                         _p(out["unsafe"]), _p(out["datax"]), _p(out["arg_link"]), _p(out["arg_obs"]), _p(out["mins"]))
    return out


def cbf_filter(n, weights, hidden, layers, datax, goal, K, u_lo, u_hi, lam, penalty, u_ref=None):
    datax = _f(datax)
    B = datax.shape[0]
    goal = _f(goal).reshape(-1, n)
    u_ref = _f(u_ref) if u_ref is not None else None
    w, K, u_lo, u_hi = _f(weights), _f(K), _f(u_lo), _f(u_hi)
    out = dict(h=np.empty(B, np.float32), J=np.empty((B, n), np.float32), u=np.empty((B, n), np.float32),
               r=np.empty(B, np.float32))
    rc = lib().oracle_cbf_filter(n, hidden, layers, _p(w), _p(datax), C.c_long(B), _p(goal), C.c_long(goal.shape[0]),
                                 _p(u_ref), _p(K), _p(u_lo), _p(u_hi), C.c_float(lam), C.c_float(penalty), _p(out["h"]),
                                 _p(out["J"]), _p(out["u"]), _p(out["r"]))
    assert rc == 0
    return out


def cbf_filter_f64(n, weights, hidden, layers, datax, goal, K, u_lo, u_hi, lam, penalty, u_ref=None):
    """float64 end-to-end truth of the filter stage + what a tolerance statement needs (oracle_cbf_filter_f64):
    h, J, u, r (float64), margin = distance to the nearest ReLU kink in pre-activation units, kappa_c / kappa_a =
    first-order conditioning of the QP solution w.r.t. c = lambda h and a = J, regime (0 inactive / 1 tight / 2 relaxed)."""
    datax = _f(datax)
    B = datax.shape[0]
    goal = _f(goal).reshape(-1, n)
    u_ref = _f(u_ref) if u_ref is not None else None
    w, K, u_lo, u_hi = _f(weights), _f(K), _f(u_lo), _f(u_hi)
    out = dict(h=np.empty(B), J=np.empty((B, n)), u=np.empty((B, n)), r=np.empty(B), margin=np.empty(B),
               kappa_c=np.empty(B), kappa_a=np.empty(B), regime=np.empty(B, np.int32))
    rc = lib().oracle_cbf_filter_f64(n, hidden, layers, _p(w), _p(datax), C.c_long(B), _p(goal), C.c_long(goal.shape[0]),
                                     _p(u_ref), _p(K), _p(u_lo), _p(u_hi), C.c_float(lam), C.c_float(penalty),
                                     _p(out["h"]), _p(out["J"]), _p(out["u"]), _p(out["r"]), _p(out["margin"]),
                                     _p(out["kappa_c"]), _p(out["kappa_a"]), _p(out["regime"]))
    assert rc == 0
    return out


def qp(a, c, u_ref, u_lo, u_hi, penalty):
    a, c, u_ref, u_lo, u_hi = _f(a), _f(c).reshape(-1), _f(u_ref), _f(u_lo), _f(u_hi)
    B, n = a.shape
    u, r = np.empty((B, n), np.float32), np.empty(B, np.float32)
    rc = lib().oracle_qp(n, _p(a), _p(c), _p(u_ref), C.c_long(B), _p(u_lo), _p(u_hi), C.c_float(penalty), _p(u), _p(r))
    assert rc == 0
    return u, r


def qp_backward(a, c, u_ref, u_lo, u_hi, penalty, grad_u, grad_r=None, want_margin=False):
    """float64 adjoint of the single-constraint QP on its active set -> (dL/da, dL/dc, dL/du_ref[, kink margin])"""
    a, c, u_ref, u_lo, u_hi, grad_u = _f(a), _f(c).reshape(-1), _f(u_ref), _f(u_lo), _f(u_hi), _f(grad_u)
    B, n = a.shape
    gr = _f(grad_r).reshape(-1) if grad_r is not None else None
    ga, gc, gt = np.empty((B, n), np.float32), np.empty(B, np.float32), np.empty((B, n), np.float32)
    mg = np.empty(B, np.float32)
    rc = lib().oracle_qp_backward(n, _p(a), _p(c), _p(u_ref), C.c_long(B), _p(u_lo), _p(u_hi), C.c_float(penalty),
                                  _p(grad_u), _p(gr) if gr is not None else None, _p(ga), _p(gc), _p(gt), _p(mg))
    assert rc == 0
    return (ga, gc, gt, mg) if want_margin else (ga, gc, gt)


def qp_multi(a, c, u_ref, u_lo, u_hi, penalty, max_iter=2000, tol=1e-10, want_iters=False):
    """float64 dual active-set solve of the S-scenario QP; iters == max_iter marks a row that did not converge"""
    a, c, u_ref, u_lo, u_hi = _f(a), _f(c), _f(u_ref), _f(u_lo), _f(u_hi)
    B, S, n = a.shape
    u, r, it = np.empty((B, n), np.float32), np.empty((B, S), np.float32), np.empty(B, np.int32)
    rc = lib().oracle_qp_multi(n, S, _p(a), _p(c), _p(u_ref), C.c_long(B), _p(u_lo), _p(u_hi), C.c_float(penalty),
                               C.c_int(max_iter), C.c_double(tol), _p(u), _p(r), _p(it))
    assert rc == 0
    return (u, r, it) if want_iters else (u, r)


def qp_multi_backward(a, c, u_ref, u_lo, u_hi, penalty, grad_u, grad_r=None, want_margin=False):
    """float64 adjoint of the S-constraint QP on its active set -> (dL/da [B,S,n], dL/dc [B,S], dL/du_ref [B,n][, margin])"""
    a, c, u_ref, u_lo, u_hi, grad_u = _f(a), _f(c), _f(u_ref), _f(u_lo), _f(u_hi), _f(grad_u)
    B, S, n = a.shape
    gr = _f(grad_r).reshape(B, S) if grad_r is not None else None
    ga, gc, gt = np.empty((B, S, n), np.float32), np.empty((B, S), np.float32), np.empty((B, n), np.float32)
    mg = np.empty(B, np.float32)
    rc = lib().oracle_qp_multi_backward(n, S, _p(a), _p(c), _p(u_ref), C.c_long(B), _p(u_lo), _p(u_hi),
                                        C.c_float(penalty), _p(grad_u), _p(gr) if gr is not None else None, _p(ga),
                                        _p(gc), _p(gt), _p(mg))
    assert rc == 0
    return (ga, gc, gt, mg) if want_margin else (ga, gc, gt)


def u_nominal(q, goal, K, u_lo, u_hi):
    """clamp(-K (q - goal), lo, hi) in float32 with the oracle's operation order (filter_state)."""
    q, goal, K = _f(q), _f(goal).reshape(-1, q.shape[1]), _f(K)
    n = q.shape[1]
    d = q - goal
    acc = np.zeros_like(q)
    for i in range(n):
        a = np.zeros(q.shape[0], np.float32)
        for k in range(n):
            a = (a.astype(np.float64) + K[i, k].astype(np.float64) * d[:, k].astype(np.float64)).astype(np.float32)
        acc[:, i] = -a
    return np.clip(acc, _f(u_lo), _f(u_hi))


def safety_filter(robot, q, boxes, spheres, floor, weights, hidden, layers, goal, K, u_lo, u_hi, thr_soft, thr_hard,
                  lam, penalty):
    q = _f(q)
    n, L, _ = robot_info(robot)
    B = q.shape[0]
    b8, sp = pack_scene(boxes, spheres)
    goal = _f(goal).reshape(-1, n)
    w, K, u_lo, u_hi = _f(weights), _f(K), _f(u_lo), _f(u_hi)
    out = dict(safe=np.empty(B, np.uint8), unsafe=np.empty(B, np.uint8), datax=np.empty((B, 2 * n + 1), np.float32),
               h=np.empty(B, np.float32), J=np.empty((B, n), np.float32), u=np.empty((B, n), np.float32),
               r=np.empty(B, np.float32))
    rc = lib().oracle_safety_filter(ROBOT_ID[robot], b8.shape[0], _p(b8), sp.shape[0], _p(sp), int(bool(floor)), hidden,
                                    layers, _p(w), _p(q), C.c_long(q.shape[1]), C.c_long(B), _p(goal),
                                    C.c_long(goal.shape[0]), _p(K), _p(u_lo), _p(u_hi), C.c_float(thr_soft),
                                    C.c_float(thr_hard), C.c_float(lam), C.c_float(penalty), _p(out["safe"]),
                                    _p(out["unsafe"]), _p(out["datax"]), _p(out["h"]), _p(out["J"]), _p(out["u"]),
                                    _p(out["r"]))
    assert rc == 0
    return out


def edge_validity(robot, q_from, q_to, n_interp, boxes, spheres=None, floor=True, thr_soft=0.02, thr_hard=0.0):
    q_from, q_to = _f(q_from), _f(q_to)
    E = q_from.shape[0]
    b8, sp = pack_scene(boxes, spheres)
    valid = np.empty(E, np.uint8)
    first_bad = np.empty(E, np.int32)
    lib().oracle_edge_validity(ROBOT_ID[robot], b8.shape[0], _p(b8), sp.shape[0], _p(sp), int(bool(floor)), _p(q_from),
                               _p(q_to), C.c_long(E), int(n_interp), C.c_float(thr_soft), C.c_float(thr_hard),
                               _p(valid), _p(first_bad))
    return valid, first_bad


def edge_validity_var(robot, q_from, q_to, n_interp, boxes, spheres=None, floor=True, thr_soft=0.02, thr_hard=0.0):
    """per-edge number of interpolation points n_interp[E] >= 0"""
    q_from, q_to = _f(q_from), _f(q_to)
    E = q_from.shape[0]
    K = np.ascontiguousarray(n_interp, dtype=np.int32).reshape(E)
    b8, sp = pack_scene(boxes, spheres)
    valid = np.empty(E, np.uint8)
    first_bad = np.empty(E, np.int32)
    lib().oracle_edge_validity_var(ROBOT_ID[robot], b8.shape[0], _p(b8), sp.shape[0], _p(sp), int(bool(floor)), _p(q_from),
                                   _p(q_to), C.c_long(E), _p(K), C.c_float(thr_soft), C.c_float(thr_hard),
                                   _p(valid), _p(first_bad))
    return valid, first_bad


def knn(queries, tree, k):
    queries, tree = _f(queries), _f(tree)
    M, n = queries.shape
    idx, dist = np.empty((M, k), np.int32), np.empty((M, k), np.float32)
    lib().oracle_knn(n, _p(queries), _p(tree), C.c_long(M), C.c_long(tree.shape[0]), int(k), _p(idx), _p(dist))
    return idx, dist


def radius_mask(queries, points, radius):
    queries, points = _f(queries), _f(points)
    M, n = queries.shape
    out = np.empty((M, points.shape[0]), np.uint8)
    lib().oracle_radius_mask(n, _p(queries), _p(points), C.c_long(M), C.c_long(points.shape[0]), C.c_float(radius), _p(out))
    return out


def rollout(robot, q0, goal, steps, ctrl_every, dt, boxes, spheres, floor, weights, hidden, layers, K, u_lo, u_hi,
            thr_soft, thr_hard, lam, penalty, stuck_lag=0, stuck_after=10, stuck_eps=0.1, want_traj=False, keep_going=False,
            stuck_mode=0):
    q0 = _f(q0)
    n, L, _ = robot_info(robot)
    T = q0.shape[0]
    goal = _f(goal).reshape(-1, n)
    b8, sp = pack_scene(boxes, spheres)
    w, K, u_lo, u_hi = _f(weights), _f(K), _f(u_lo), _f(u_hi)
    qf = np.empty((T, n), np.float32)
    alive = np.empty(T, np.uint8)
    done = np.empty(T, np.int32)
    traj = np.zeros((T, steps, n), np.float32) if (want_traj or stuck_lag > 0) else None
    rc = lib().oracle_rollout(ROBOT_ID[robot], b8.shape[0], _p(b8), sp.shape[0], _p(sp), int(bool(floor)), hidden,
                              layers, _p(w), _p(q0), _p(goal), C.c_long(goal.shape[0]), C.c_long(T), int(steps),
                              int(ctrl_every), C.c_float(dt), int(stuck_lag), int(stuck_after), C.c_float(stuck_eps),
                              int(bool(keep_going)), _p(K), _p(u_lo), _p(u_hi), C.c_float(thr_soft),
                              C.c_float(thr_hard), C.c_float(lam), C.c_float(penalty), _p(qf), _p(alive), _p(done),
                              _p(traj), int(stuck_mode))
    assert rc == 0
    return qf, alive, done, traj


def rollout_segments(robot, q0, goal, n_seg, seg_len, ctrl_every, dt, boxes, spheres, floor, weights, hidden, layers, K, u_lo,
                     u_hi, thr_soft, thr_hard, lam, penalty, stuck_lag=9, stuck_after=10, stuck_eps=3e-2):
    """n_seg consecutive batched steer segments of one batch (batch_tsa.py:154-170 over :191-231)
    -> seg_q [T, n_seg, n], seg_alive [T, n_seg], seg_done [T, n_seg], traj [T, n_seg * seg_len, n]"""
    q0 = _f(q0)
    n, L, _ = robot_info(robot)
    T = q0.shape[0]
    goal = _f(goal).reshape(-1, n)
    b8, sp = pack_scene(boxes, spheres)
    w, K, u_lo, u_hi = _f(weights), _f(K), _f(u_lo), _f(u_hi)
    seg_q = np.empty((T, n_seg, n), np.float32)
    seg_alive = np.empty((T, n_seg), np.uint8)
    seg_done = np.empty((T, n_seg), np.int32)
    traj = np.zeros((T, n_seg * seg_len, n), np.float32)
    rc = lib().oracle_rollout_segments(ROBOT_ID[robot], b8.shape[0], _p(b8), sp.shape[0], _p(sp), int(bool(floor)), hidden,
                                       layers, _p(w), _p(q0), _p(goal), C.c_long(goal.shape[0]), C.c_long(T), int(n_seg),
                                       int(seg_len), int(ctrl_every), C.c_float(dt), int(stuck_lag), int(stuck_after),
                                       C.c_float(stuck_eps), _p(K), _p(u_lo), _p(u_hi), C.c_float(thr_soft),
                                       C.c_float(thr_hard), C.c_float(lam), C.c_float(penalty), _p(seg_q), _p(seg_alive),
                                       _p(seg_done), _p(traj))
    assert rc == 0
    return seg_q, seg_alive, seg_done, traj


def sample_states(robot, seed, start_row, B):
    n, _, _ = robot_info(robot)
    q = np.empty((B, n), np.float32)
    lib().oracle_sample_states(ROBOT_ID[robot], C.c_uint32(seed), C.c_uint64(start_row), C.c_long(B), _p(q))
    return q
