"""CPU suite: the C-ABI library loads, exports every symbol include/cbfrrt.h declares, and validates arguments
before touching CUDA (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cbfrrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cbfrrt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cbfrrt import _lib
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/cbfrrt.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in cbfrrt/_lib.py"
    assert _lib.lib.cbfrrt_abi_version() == 2


def test_robot_info_and_status_strings():
    from cbfrrt import _lib
    assert _lib.robot_info("panda") == (7, 12, 32)
    assert _lib.robot_info("magician") == (4, 4, 24)
    n = C.c_int32()
    assert _lib.lib.cbfrrt_robot_info(9, C.byref(n), C.byref(n), C.byref(n)) == -1
    assert _lib.lib.cbfrrt_status_string(-3) == b"unsupported network shape or scene size"


def test_argument_validation_happens_before_cuda():
    from cbfrrt import _lib
    lib = _lib.lib
    sc = _lib.Scene(0, 0, 1, 0, None, None, None)
    assert lib.cbfrrt_grid_bytes() == 44 * 44 * 44 * 24 * 2 + 88 * 88 * 88 * 4   # candidate records + distance-bound field
    assert lib.cbfrrt_grid_build(C.byref(sc), None, None) == -2
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p)
    assert lib.cbfrrt_fk(5, p, 7, 1, p, p, None) == -1                                   # bad robot
    assert lib.cbfrrt_fk(0, p, 7, -1, p, p, None) == -2                                  # negative batch
    assert lib.cbfrrt_fk(0, p, 3, 1, p, p, None) == -2                                   # stride < n
    assert lib.cbfrrt_collide(0, C.byref(sc), None, 7, 1, 0.02, 0.0, *([None] * 6), None) == -2
    big = _lib.Scene(2000, 0, 1, 0, p, None, None)
    assert lib.cbfrrt_collide(0, C.byref(big), p, 7, 1, 0.02, 0.0, *([None] * 6), None) == -3  # > MAX_OBSTACLES
    mis = _lib.Scene(1, 0, 1, 0, C.c_void_p(p.value + 4), None, None)
    assert lib.cbfrrt_collide(0, C.byref(mis), p, 7, 1, 0.02, 0.0, *([None] * 6), None) == -5  # alignment
    net = _lib.CbfNet(8, 1024, 2, 0, p)
    fp = _lib.FilterParams(p, 1, p, p, p, 0.1, 5000.0, None)
    assert lib.cbfrrt_cbf_filter(7, C.byref(net), p, 1, C.byref(fp), p, p, p, p, None) == -3   # (layers + 2) * hidden > 1536
    net = _lib.CbfNet(8, 64, 3, 0, p)                                                      # generic shape: fine for the
    rp = _lib.RolloutParams(10, 4, 0.01, 0, 10, 0.1, 0, 0)                                 # filter, not for the fused rollout
    assert lib.cbfrrt_rollout(0, C.byref(sc), C.byref(net), p, 1, C.byref(fp), C.byref(rp), 0.02, 0.0, p, p, p, None, None) == -3
    assert lib.cbfrrt_safety_filter(0, C.byref(sc), C.byref(net), p, 7, 1, C.byref(fp), 0.02, 0.0, p, p, None, p, p, p, p, None) == -3  # needs datax
    net = _lib.CbfNet(8, 48, 2, 0, p)
    fp = _lib.FilterParams(p, 3, p, p, p, 0.1, 5000.0, None)
    assert lib.cbfrrt_cbf_filter(7, C.byref(net), p, 5, C.byref(fp), p, p, p, p, None) == -2   # goal_rows not in {1,B}
    assert lib.cbfrrt_qp(9, p, p, p, 1, p, p, 1.0, p, p, None) == -3
    assert lib.cbfrrt_qp_backward(9, p, p, p, 1, p, p, 1.0, p, p, p, p, p, None) == -3
    assert lib.cbfrrt_qp_backward(7, None, p, p, 1, p, p, 1.0, p, p, p, p, p, None) == -2
    assert lib.cbfrrt_qp_backward(7, None, None, None, 0, p, p, 1.0, None, None, None, None, None, None) == 0
    assert lib.cbfrrt_qp_multi_backward(7, 5, p, p, p, 1, p, p, 1.0, 10, 1e-9, p, p, p, p, p, None) == -3   # > 4 constraints
    assert lib.cbfrrt_qp_multi_backward(7, 2, p, p, p, 1, p, p, 1.0, 0, 1e-9, p, p, p, p, p, None) == -2    # max_iter < 1
    # knn workspace: 16 M bytes for k = 1, + 8 M ceil(N / 1024) k for the two-launch k-nearest path
    assert lib.cbfrrt_knn_workspace_bytes(50, 100000, 1) == 800
    assert lib.cbfrrt_knn_workspace_bytes(1, 100000, 50) == 16 + 8 * 98 * 50
    assert lib.cbfrrt_knn_workspace_bytes(3, 40, 64) == 48 + 8 * 3 * 64
    assert lib.cbfrrt_knn_workspace_bytes(2, 10, 5000) == 32 and lib.cbfrrt_knn_workspace_bytes(-1, 1, 1) == -1
    # B == 0 is a valid no-op everywhere
    assert lib.cbfrrt_fk(0, p, 7, 0, p, p, None) == 0
    assert lib.cbfrrt_collide(0, C.byref(sc), p, 7, 0, 0.02, 0.0, *([None] * 6), None) == 0


def test_ops_refuse_cpu_tensors():
    """CUDA-only registration: no CPU kernel exists, so the dispatcher must raise (no silent fallback)."""
    import torch
    from cbfrrt import ops
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.fk(torch.zeros(2, 7), 0)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.masks(torch.zeros(2, 7), torch.zeros(0, 8), torch.zeros(0, 4), 1, 0.02, 0.0, 0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cbf-inc-rrt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "c_oracle" not in txt, os.path.join(dp, f)


def test_fused_exchange_stores_are_multimem_in_the_ptx(tmp_path):
    """The multicast form of the fused exchange must access the NVSwitch multicast address with multimem.* (PTX ISA).
    ptxas lowers multimem.st to STG.E.STRONG.SYS on sm_100a, so the SASS shows no distinct mnemonic: the evidence is the
    PTX of the store helpers of csrc/kernels.cuh, compiled here."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "mm.cu"
    src.write_text('#include "%s"\n__global__ void probe(float* p, float4 v) {\n'
                   '    cbfrrt::multimem_st_v4(p, v.x, v.y, v.z, v.w);\n    cbfrrt::multimem_st_f32(p + 4, v.x);\n'
                   '    cbfrrt::flush_block<true>(p, threadIdx.x, p + 8, 32);\n}\n'
                   % os.path.join(ROOT, "cbf-inc-rrt_b200", "csrc", "kernels.cuh"))
    ptx = tmp_path / "mm.ptx"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=compute_100a", "-std=c++17", "--expt-relaxed-constexpr",
                           "-diag-suppress", "20091", "-ptx", "-o", str(ptx), str(src)])
    text = ptx.read_text()
    assert text.count("multimem.st.relaxed.sys.global.v4.f32") >= 2 and "multimem.st.relaxed.sys.global.f32" in text
