"""ctypes binding of libcbfrrt.so (include/cbfrrt.h).  There is NO fallback: if the CUDA library has not been
built (python __graft_entry__.py / cbf-inc-rrt_b200/csrc/build.sh) importing this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libcbfrrt.so")

ROBOT_ID = {"panda": 0, "magician": 1}

STATUS = {0: "ok", -1: "unknown robot id", -2: "bad shape or missing pointer",
          -3: "unsupported network shape or scene size", -4: "CUDA launch/runtime failure",
          -5: "scene/weight buffer not 16-byte aligned"}


class Scene(C.Structure):
    _fields_ = [("n_boxes", C.c_int32), ("n_spheres", C.c_int32), ("floor", C.c_int32), ("_pad", C.c_int32),
                ("boxes", C.c_void_p), ("spheres", C.c_void_p), ("grid", C.c_void_p), ("stats", C.c_void_p)]


class HostScene(C.Structure):
    _fields_ = [("n_boxes", C.c_int32), ("n_spheres", C.c_int32), ("floor", C.c_int32), ("_pad", C.c_int32),
                ("boxes", C.c_void_p), ("spheres", C.c_void_p)]


class CbfNet(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("_pad", C.c_int32),
                ("weights", C.c_void_p)]


class FilterParams(C.Structure):
    _fields_ = [("goal", C.c_void_p), ("goal_rows", C.c_int64), ("K", C.c_void_p), ("u_lo", C.c_void_p),
                ("u_hi", C.c_void_p), ("cbf_lambda", C.c_float), ("relaxation_penalty", C.c_float),
                ("u_ref", C.c_void_p)]


class PeerOut(C.Structure):
    _fields_ = [("n_peers", C.c_int32), ("multicast", C.c_int32), ("row_offset", C.c_int64), ("valid", C.c_void_p * 8),
                ("u", C.c_void_p * 8)]


class LidarNet(C.Structure):
    _fields_ = ([(k, C.c_int32) for k in ("n", "n_sensors", "point_dims", "per_feature_dim", "feature_dim", "hidden", "layers", "_pad")]
                + [(k, C.c_void_p) for k in ("w1t", "w2t", "w3t", "bias", "fc2_w", "fc2_b", "fc3_w", "fc3_b", "e1_w", "e1_b", "e2_w",
                                             "e2_b", "e4_w", "e4_b", "h_in_w", "h_in_b", "h_hid_w", "h_hid_b", "h_out_w", "h_out_b")])


class RolloutParams(C.Structure):
    _fields_ = [("steps", C.c_int32), ("ctrl_every", C.c_int32), ("dt", C.c_float), ("stuck_lag", C.c_int32),
                ("stuck_after", C.c_int32), ("stuck_eps", C.c_float), ("keep_going", C.c_int32),
                ("stuck_mode", C.c_int32)]


# every symbol declared in include/cbfrrt.h: (restype, argtypes)
_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
SIGNATURES = {
    "cbfrrt_abi_version": (C.c_int, []),
    "cbfrrt_status_string": (C.c_char_p, [_I]),
    "cbfrrt_last_cuda_error": (C.c_int, []),
    "cbfrrt_launch_count": (C.c_int64, []),
    "cbfrrt_set_mlp_path": (C.c_int, [_I]),
    "cbfrrt_get_mlp_path": (C.c_int, []),
    "cbfrrt_robot_info": (C.c_int, [_I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "cbfrrt_grid_bytes": (C.c_int64, []),
    "cbfrrt_grid_build": (C.c_int, [C.POINTER(Scene), _P, _P]),
    "cbfrrt_fk": (C.c_int, [_I, _P, _L, _L, _P, _P, _P]),
    "cbfrrt_collide": (C.c_int, [_I, C.POINTER(Scene), _P, _L, _L, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "cbfrrt_cbf_filter": (C.c_int, [_I, C.POINTER(CbfNet), _P, _L, C.POINTER(FilterParams), _P, _P, _P, _P, _P]),
    "cbfrrt_qp": (C.c_int, [_I, _P, _P, _P, _L, _P, _P, _F, _P, _P, _P]),
    "cbfrrt_qp_backward": (C.c_int, [_I, _P, _P, _P, _L, _P, _P, _F, _P, _P, _P, _P, _P, _P]),
    "cbfrrt_knn_workspace_bytes": (C.c_int64, [_L, _L, C.c_int32]),
    "cbfrrt_knn": (C.c_int, [_I, _P, _P, _L, _L, C.c_int32, _P, _P, _P, _P]),
    "cbfrrt_radius_mask": (C.c_int, [_I, _P, _P, _L, _L, _F, _P, _P]),
    "cbfrrt_qp_multi": (C.c_int, [_I, _I, _P, _P, _P, _L, _P, _P, _F, _I, _F, _P, _P, _P, _P]),
    "cbfrrt_qp_multi_backward": (C.c_int, [_I, _I, _P, _P, _P, _L, _P, _P, _F, _I, _F, _P, _P, _P, _P, _P, _P]),
    "cbfrrt_safety_filter": (C.c_int, [_I, C.POINTER(Scene), C.POINTER(CbfNet), _P, _L, _L, C.POINTER(FilterParams),
                                       _F, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cbfrrt_safety_filter_scatter": (C.c_int, [_I, C.POINTER(Scene), C.POINTER(CbfNet), _P, _L, _L,
                                               C.POINTER(FilterParams), _F, _F, C.POINTER(PeerOut), _P]),
    "cbfrrt_edge_validity": (C.c_int, [_I, C.POINTER(Scene), _P, _P, _L, C.c_int32, _F, _F, _P, _P, _P]),
    "cbfrrt_edge_validity_var": (C.c_int, [_I, C.POINTER(Scene), _P, _P, _L, _P, _L, _F, _F, _P, _P, _P]),
    "cbfrrt_rollout": (C.c_int, [_I, C.POINTER(Scene), C.POINTER(CbfNet), _P, _L, C.POINTER(FilterParams),
                                 C.POINTER(RolloutParams), _F, _F, _P, _P, _P, _P, _P]),
    "cbfrrt_rollout_segments": (C.c_int, [_I, C.POINTER(Scene), C.POINTER(CbfNet), _P, _L, C.POINTER(FilterParams),
                                          C.POINTER(RolloutParams), C.c_int32, _F, _F, _P, _P, _P, _P, _P]),
    "cbfrrt_fma_probe": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "cbfrrt_lidar_workspace_bytes": (C.c_int64, [C.POINTER(LidarNet), _L, C.c_int32]),
    "cbfrrt_lidar_cbf": (C.c_int, [C.POINTER(LidarNet), _P, _P, _L, _P, _P, _P, _P, C.c_int32, _L, _F, _P, _P, _P, _P]),
    "cbfrrt_sample_states": (C.c_int, [_I, C.c_uint32, C.c_uint64, _L, _P, _P]),
    "cbfrrt_safety_filter_host": (C.c_int, [_I, _I, C.POINTER(HostScene), C.c_int32, C.c_int32, _P, _L, _P, _L, _P, _L,
                                            _P, _P, _P, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "cbfrrt_rollout_segments_host": (C.c_int, [_I, _I, C.POINTER(HostScene), C.c_int32, C.c_int32, _P, _L, _P, _L, _P, _P, _P, _P,
                                               _F, _F, C.POINTER(RolloutParams), C.c_int32, _F, _F, _P, _P, _P, _P]),
    "cbfrrt_cbf_filter_scatter": (C.c_int, [_I, C.POINTER(CbfNet), _P, _P, _L, C.POINTER(FilterParams), C.POINTER(PeerOut), _P]),
    "cbfrrt_set_workspace": (C.c_int, [_P, _L]),
    "cbfrrt_set_two_pass": (C.c_int, [_I]),
    "cbfrrt_knn_host": (C.c_int, [_I, _I, _P, _P, _L, _L, C.c_int32, _P, _P]),
    "cbfrrt_edge_validity_host": (C.c_int, [_I, _I, C.POINTER(HostScene), _P, _P, _L, _P, _F, _F, _P, _P]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA extension has not been built (run `python __graft_entry__.py` or "
        f"cbf-inc-rrt_b200/csrc/build.sh).  cbfrrt has no CPU fallback.")

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

ABI_VERSION = lib.cbfrrt_abi_version()


class CbfrrtError(RuntimeError):
    pass


def check(status, what):
    if status != 0:
        extra = f" (cudaError {lib.cbfrrt_last_cuda_error()})" if status == -4 else ""
        raise CbfrrtError(f"{what} failed: {STATUS.get(status, status)}{extra}")


def robot_info(robot):
    n, l, s = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib.cbfrrt_robot_info(ROBOT_ID[robot], C.byref(n), C.byref(l), C.byref(s)), "cbfrrt_robot_info")
    return n.value, l.value, s.value
