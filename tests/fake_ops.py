"""TEST-ONLY stand-ins for the CUDA ops, backed by the C oracle, operating on CPU tensors.

They let the CPU suite exercise the HOST logic of the reference-shaped API (shapes, dtypes, devices, return
conventions, control flow of steer / edge_checking, sharding) without a GPU.  The product never sees them."""
import numpy as np
import torch

from oracle import c_oracle

NAME = {0: "panda", 1: "magician"}


def _np(t):
    return t.detach().cpu().numpy()


def _scene(boxes8, spheres):
    b = _np(boxes8)[:, :6] if boxes8.numel() else None
    s = _np(spheres) if spheres.numel() else None
    return b, s


def install(monkeypatch):
    from cbfrrt import ops

    def fk(q, robot_id):
        pos, rot = c_oracle.fk(NAME[robot_id], _np(q))
        return torch.from_numpy(pos), torch.from_numpy(rot)

    def collide(q, boxes, spheres, floor, grid, thr_soft, thr_hard, robot_id):
        b, s = _scene(boxes, spheres)
        o = c_oracle.collide(NAME[robot_id], _np(q)[:, :ops.ROBOT_N[robot_id]], b, s, bool(floor), thr_soft, thr_hard)
        return tuple(torch.from_numpy(o[k]) for k in ("safe", "unsafe", "datax", "arg_link", "arg_obs", "mins"))

    def observe(*a):
        return collide(*a)[:3]

    def masks(*a):
        return collide(*a)[:2]

    def cbf_filter(datax, goal, weights, K, u_lo, u_hi, lam, penalty, hidden, layers, u_ref):
        n = (datax.shape[1] - 1) // 2
        o = c_oracle.cbf_filter(n, _np(weights), hidden, layers, _np(datax), _np(goal), _np(K), _np(u_lo), _np(u_hi), lam,
                                penalty, u_ref=_np(u_ref) if u_ref.numel() else None)
        return tuple(torch.from_numpy(o[k]) for k in ("h", "J", "u", "r"))

    def qp(a, c, u_ref, u_lo, u_hi, penalty):
        u, r = c_oracle.qp(_np(a), _np(c), _np(u_ref), _np(u_lo), _np(u_hi), penalty)
        return torch.from_numpy(u), torch.from_numpy(r)

    def qp_backward(a, c, u_ref, u_lo, u_hi, penalty, grad_u, grad_r):
        ga, gc, gt = c_oracle.qp_backward(_np(a), _np(c), _np(u_ref), _np(u_lo), _np(u_hi), penalty, _np(grad_u),
                                          _np(grad_r) if grad_r.numel() else None)
        return torch.from_numpy(ga), torch.from_numpy(gc), torch.from_numpy(gt)

    def qp_multi(a, c, u_ref, u_lo, u_hi, penalty, max_sweeps, tol):
        u, r = c_oracle.qp_multi(_np(a), _np(c), _np(u_ref), _np(u_lo), _np(u_hi), penalty)
        return torch.from_numpy(u), torch.from_numpy(r), torch.zeros(a.shape[0], dtype=torch.int32)

    def qp_multi_backward(a, c, u_ref, u_lo, u_hi, penalty, max_sweeps, tol, grad_u, grad_r):
        ga, gc, gt = c_oracle.qp_multi_backward(_np(a), _np(c), _np(u_ref), _np(u_lo), _np(u_hi), penalty, _np(grad_u),
                                                _np(grad_r) if grad_r.numel() else None)
        return torch.from_numpy(ga), torch.from_numpy(gc), torch.from_numpy(gt)

    def edge_validity(q_from, q_to, n_interp, boxes, spheres, floor, grid, thr_soft, thr_hard, robot_id):
        b, s = _scene(boxes, spheres)
        v, f = c_oracle.edge_validity(NAME[robot_id], _np(q_from), _np(q_to), n_interp, b, s, bool(floor), thr_soft, thr_hard)
        return torch.from_numpy(v), torch.from_numpy(f)

    def edge_validity_var(q_from, q_to, n_interp, boxes, spheres, floor, grid, thr_soft, thr_hard, robot_id):
        b, s = _scene(boxes, spheres)
        v, f = c_oracle.edge_validity_var(NAME[robot_id], _np(q_from), _np(q_to), _np(n_interp), b, s, bool(floor), thr_soft,
                                          thr_hard)
        return torch.from_numpy(v), torch.from_numpy(f)

    def knn(queries, tree, k):
        i, d = c_oracle.knn(_np(queries), _np(tree), k)
        return torch.from_numpy(i), torch.from_numpy(d)

    def radius_mask(queries, points, radius):
        return torch.from_numpy(c_oracle.radius_mask(_np(queries), _np(points), radius))

    def rollout(q0, goal, steps, ctrl_every, dt, stuck_lag, stuck_after, stuck_eps, want_traj, keep_going, boxes, spheres, floor, grid,
                weights, K, u_lo, u_hi, lam, penalty, thr_soft, thr_hard, robot_id, hidden, layers, stuck_mode=0):
        b, s = _scene(boxes, spheres)
        qf, alive, done, traj = c_oracle.rollout(NAME[robot_id], _np(q0), _np(goal), steps, ctrl_every, dt, b, s, bool(floor),
                                                 _np(weights), hidden, layers, _np(K), _np(u_lo), _np(u_hi), thr_soft,
                                                 thr_hard, lam, penalty, stuck_lag=stuck_lag, stuck_after=stuck_after,
                                                 stuck_eps=stuck_eps, want_traj=want_traj, keep_going=keep_going, stuck_mode=stuck_mode)
        return (torch.from_numpy(qf), torch.from_numpy(alive), torch.from_numpy(done),
                torch.from_numpy(traj) if traj is not None else torch.empty(0))

    def rollout_segments(q0, goal, n_segments, seg_len, ctrl_every, dt, stuck_lag, stuck_after, stuck_eps, boxes, spheres, floor, grid,
                         weights, K, u_lo, u_hi, lam, penalty, thr_soft, thr_hard, robot_id, hidden, layers):
        b, s = _scene(boxes, spheres)
        sq, sa, sd, tr = c_oracle.rollout_segments(NAME[robot_id], _np(q0), _np(goal), n_segments, seg_len, ctrl_every, dt, b, s,
                                                   bool(floor), _np(weights), hidden, layers, _np(K), _np(u_lo), _np(u_hi), thr_soft,
                                                   thr_hard, lam, penalty, stuck_lag=stuck_lag, stuck_after=stuck_after,
                                                   stuck_eps=stuck_eps)
        return torch.from_numpy(sq), torch.from_numpy(sa), torch.from_numpy(sd), torch.from_numpy(tr)

    def rollout_segments_host(robot, q0, goal, n_segments, seg_len, ctrl_every, dt, stuck_lag, stuck_after, stuck_eps, boxes6, spheres,
                              floor, weights, hidden, layers, K, u_lo, u_hi, lam, penalty, thr_soft, thr_hard, device=0):
        f = lambda a: np.ascontiguousarray(a, np.float32)
        sq, sa, sd, tr = c_oracle.rollout_segments(robot, f(q0), f(goal), n_segments, seg_len, ctrl_every, dt, f(boxes6).reshape(-1, 6),
                                                   None if spheres is None or len(spheres) == 0 else f(spheres), bool(floor), f(weights),
                                                   hidden, layers, f(K), f(u_lo), f(u_hi), thr_soft, thr_hard, lam, penalty,
                                                   stuck_lag=stuck_lag, stuck_after=stuck_after, stuck_eps=stuck_eps)
        return sq, sa.astype(bool), sd, tr

    def edge_validity_host(robot, q_from, q_to, n_interp, boxes6, spheres, floor, thr_soft, thr_hard, device=0):
        f = lambda a: np.ascontiguousarray(a, np.float32)
        v, fb = c_oracle.edge_validity_var(robot, f(q_from), f(q_to), np.asarray(n_interp, np.int32), f(boxes6).reshape(-1, 6),
                                           None if spheres is None or len(spheres) == 0 else f(spheres), bool(floor), thr_soft, thr_hard)
        return v.astype(bool), fb

    def knn_host(queries, tree, k, device=0):
        return c_oracle.knn(np.ascontiguousarray(queries, np.float32), np.ascontiguousarray(tree, np.float32), k)

    def valid_control(q, goal, boxes, spheres, floor, grid, weights, K, u_lo, u_hi, lam, penalty, thr_soft, thr_hard, robot_id,
                      hidden, layers):
        b, s = _scene(boxes, spheres)
        o = c_oracle.safety_filter(NAME[robot_id], _np(q), b, s, bool(floor), _np(weights), hidden, layers, _np(goal), _np(K),
                                   _np(u_lo), _np(u_hi), thr_soft, thr_hard, lam, penalty)
        return torch.from_numpy(o["safe"]), torch.from_numpy(o["u"])

    for name, fn in dict(fk=fk, collide=collide, observe=observe, masks=masks, cbf_filter=cbf_filter, qp=qp, qp_backward=qp_backward, qp_multi=qp_multi, qp_multi_backward=qp_multi_backward,
                         edge_validity=edge_validity, edge_validity_var=edge_validity_var, knn=knn, radius_mask=radius_mask, rollout=rollout, rollout_segments=rollout_segments, rollout_segments_host=rollout_segments_host,
                         knn_host=knn_host, edge_validity_host=edge_validity_host, valid_control=valid_control).items():
        monkeypatch.setattr(ops, name, fn)
