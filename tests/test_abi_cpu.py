"""CPU suite, part 2: the C-ABI library loads without a GPU and exports exactly what include/quiver_b200.h declares;
argument validation that needs no device; errors are reported, never fatal."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "quiver_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"QV_API\s+[\w\s\*]+?\b(qv_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _header_symbols()
    for must in ("qv_gather", "qv_sample_count", "qv_sample_fill", "qv_reindex", "qv_khop", "qv_sampler_create",
                 "qv_init_p2p", "qv_can_device_access_peer", "qv_host_register", "qv_ipc_get_handle",
                 "qv_ipc_open_handle", "qv_cal_neighbor_prob"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from torch_quiver import _lib
    assert os.path.exists(_lib.LIB_PATH)
    dyn = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (qv_\w+)", dyn))
    declared = set(_header_symbols())
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert exported <= declared, f"exported but not declared: {sorted(exported - declared)}"
    assert set(_lib.PROTOTYPES) == declared  # the Python adapter binds exactly the header's surface
    assert _lib.lib.qv_abi_version() == 2


def test_library_is_sm100a_only():
    from torch_quiver import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_struct_layout_matches_header():
    from torch_quiver._lib import ShardTable
    # int32 n_shards, int32 reserved, int64[17], ptr[16], int64[16], int32[16]
    assert ctypes.sizeof(ShardTable) == 8 + 17 * 8 + 16 * 8 + 16 * 8 + 16 * 4
    assert ShardTable.row_begin.offset == 8 and ShardTable.ptr.offset == 8 + 17 * 8


def test_argument_errors_are_reported_not_fatal():
    from torch_quiver import _lib
    lib = _lib.lib
    rc = lib.qv_khop_bounds(16, (ctypes.c_int64 * 2)(5, -1), 2, (ctypes.c_int64 * 3)(), (ctypes.c_int64 * 2)())
    assert rc == _lib.QV_ERR_UNSUPPORTED and b"per-hop" in lib.qv_last_error()
    bn, be = (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 3)()
    assert lib.qv_khop_bounds(1024, (ctypes.c_int64 * 3)(15, 10, 5), 3, bn, be) == 0
    assert list(bn) == [1024, 1024 * 16, 1024 * 16 * 11, 1024 * 16 * 11 * 6]
    assert list(be) == [1024 * 15, 1024 * 16 * 10, 1024 * 16 * 11 * 5]
    t = _lib.ShardTable()
    t.n_shards = 0
    assert lib.qv_gather(ctypes.byref(t), None, None, 1, 16, None, 0, None) == _lib.QV_ERR_INVALID
    with pytest.raises(_lib.QuiverError):
        _lib.check(lib.qv_sampler_create(0, None, 0, None, 0, ctypes.byref(ctypes.c_void_p())))


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less machines")
    import torch_quiver
    with pytest.raises(RuntimeError):
        torch_quiver.can_device_access_peer(0, 1)
    with pytest.raises(RuntimeError):
        torch_quiver.device_quiver_from_csr_array(torch.tensor([0, 1]), torch.tensor([0]), None, 0, True)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torch-quiver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_xorwow_jump_ahead_equals_sequential_stepping(tmp_path):
    """The chain splitting of mega rows positions generators with xorwow_jump (GF(2) matrices A^(2^i), qv_xorwow.cuh).
    The function is __host__ __device__: a host build of tests/native/xorwow_jump_test.cu checks it against stepping the
    generator draw by draw (offsets 0 .. 1,000,003 on several seeds)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    csrc = os.path.join(ROOT, "torch-quiver_b200", "csrc")
    exe = str(tmp_path / "jump_test")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"),
                           "-I", csrc, os.path.join(ROOT, "tests", "native", "xorwow_jump_test.cu"),
                           os.path.join(csrc, "qv_xorwow.cu"), os.path.join(csrc, "qv_runtime.cu"), "-o", exe])
    out = subprocess.check_output([exe]).decode()
    assert "jump ok" in out


def test_compiled_call_path_loads_and_can_be_switched_off():
    """The ctypes package routes its per-step calls through the compiled adapter when it is built (build() builds it; this
    test does if it is missing): the module loads without a GPU, reports the library's ABI version and exports the two call
    shims; the A-B switch leaves the ctypes path in charge (fresh interpreters: the choice is made at import)."""
    import sys
    from conftest import PKG
    subprocess.check_call([sys.executable, os.path.join(PKG, "csrc", "pybind", "build.py")], stdout=subprocess.DEVNULL)
    code = ("import torch_quiver as qv; c = qv._compiled; "
            "print(c is not None and c.abi_version() == qv.lib.qv_abi_version() == 2 and "
            "all(hasattr(c, n) for n in ('khop_raw', 'gather_raw', 'Quiver', 'ShardTensor', 'ShardTensorItem')))")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(sys.path))
    env.pop("QUIVER_B200_COMPILED_CALLS", None)
    assert subprocess.check_output([sys.executable, "-c", code], env=env, text=True).strip() == "True"
    code = "import torch_quiver as qv; print(qv._compiled is None)"
    env["QUIVER_B200_COMPILED_CALLS"] = "0"
    assert subprocess.check_output([sys.executable, "-c", code], env=env, text=True).strip() == "True"
