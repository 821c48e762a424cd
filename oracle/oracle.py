"""Python face of the CPU oracle (TEST INFRASTRUCTURE -- see qv_oracle.c's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
numpy arrays in, numpy arrays out; every function is a direct call into libqv_oracle.so (plain C) except `khop`, which
composes them exactly as GraphSageSampler.sample does (srcs/python/quiver/pyg/sage_sampler.py:118-147).
"""
import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "libqv_oracle.so")
_REF_DIR = os.path.join(HERE, "_ref")


def build():
    """Compile the C restatement (and the cuRAND host probe); cheap, idempotent."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])


def _load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(HERE, "qv_oracle.c")):
        build()
    lib = ctypes.CDLL(_LIB)
    i64, u64, u32, vp = ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p
    lib.qo_xorwow_next.restype = u32
    lib.qo_xorwow_next.argtypes = [vp]
    lib.qo_xorwow_init.restype = None
    lib.qo_xorwow_init.argtypes = [u64, u64, u64, vp]
    lib.qo_xorwow_seq_matrix.restype = None
    lib.qo_xorwow_seq_matrix.argtypes = [u64, vp]
    lib.qo_sample_counts.restype = i64
    lib.qo_sample_counts.argtypes = [vp, i64, i64, vp, i64, i64, vp, vp]
    lib.qo_sample_neighbor_gpu.restype = None
    lib.qo_sample_neighbor_gpu.argtypes = [u64, i64, i64, vp, vp, vp, vp, vp]
    lib.qo_sample_neighbor_gpu_pos.restype = None
    lib.qo_sample_neighbor_gpu_pos.argtypes = [u64, i64, i64, vp, vp, vp, vp, vp, vp]
    lib.qo_sample_neighbor_gpu_split.restype = None
    lib.qo_sample_neighbor_gpu_split.argtypes = [u64, i64, i64, vp, vp, vp, vp, vp, i64, i64]
    lib.qo_reindex.restype = i64
    lib.qo_reindex.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp]
    lib.qo_gather.restype = None
    lib.qo_gather.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp, i64, i64, vp]
    lib.qo_cal_next.restype = None
    lib.qo_cal_next.argtypes = [vp, vp, i64, ctypes.c_int, vp, vp]
    lib.qo_validate_sample.restype = i64
    lib.qo_validate_sample.argtypes = [vp, vp, vp, i64, i64, vp, vp, i64]
    return lib


_lib = _load()


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


class Xorwow(ctypes.Structure):
    _fields_ = [("d", ctypes.c_uint32), ("v", ctypes.c_uint32 * 5)]


def xorwow_stream(seed, subseq, n, offset=0):
    """(state after curand_init(seed, subseq, offset) as [d, v0..v4], first n draws)."""
    st = Xorwow()
    _lib.qo_xorwow_init(seed, subseq, offset, ctypes.byref(st))
    state = [st.d] + list(st.v)
    return state, [int(_lib.qo_xorwow_next(ctypes.byref(st))) for _ in range(n)]


def xorwow_seq_matrix(nseq):
    out = np.zeros(160 * 5, dtype=np.uint32)
    _lib.qo_xorwow_seq_matrix(nseq, _p(out))
    return out.reshape(160, 5)


def sample_counts(indptr, seeds, k):
    indptr, seeds = _i64(indptr), _i64(seeds)
    S = seeds.shape[0]
    counts, out_ptr = np.zeros(S, np.int64), np.zeros(S, np.int64)
    n_nodes = indptr.shape[0] - 1
    tot = _lib.qo_sample_counts(_p(indptr), n_nodes, int(indptr[-1]), _p(seeds), S, int(k), _p(counts), _p(out_ptr))
    return counts, out_ptr, int(tot)


def sample_neighbor(indptr, indices, seeds, k, rand_seed=0):
    """Quiver.sample_neighbor on the reference GPU path: (neighbors, counts), bit-exact for generator seed rand_seed."""
    indptr, indices, seeds = _i64(indptr), _i64(indices), _i64(seeds)
    if k < 0:
        k = max(int(indptr.shape[0]) - 1, int(np.max(np.diff(indptr), initial=0)))  # sage_sampler.py:90
    counts, out_ptr, tot = sample_counts(indptr, seeds, k)
    out = np.zeros(tot, np.int64)
    _lib.qo_sample_neighbor_gpu(int(rand_seed), int(k), seeds.shape[0], _p(seeds), _p(indptr), _p(indices), _p(out_ptr),
                                _p(out))
    return out, counts


def sample_neighbor_pos(indptr, indices, seeds, k, rand_seed=0):
    """sample_neighbor plus the CSR position of every pick: (neighbors, counts, positions)."""
    indptr, indices, seeds = _i64(indptr), _i64(indices), _i64(seeds)
    if k < 0:
        k = max(int(indptr.shape[0]) - 1, int(np.max(np.diff(indptr), initial=0)))
    counts, out_ptr, tot = sample_counts(indptr, seeds, k)
    out, pos = np.zeros(tot, np.int64), np.zeros(tot, np.int64)
    _lib.qo_sample_neighbor_gpu_pos(int(rand_seed), int(k), seeds.shape[0], _p(seeds), _p(indptr), _p(indices),
                                    _p(out_ptr), _p(out), _p(pos))
    return out, counts, pos


def sample_neighbor_split(indptr, indices, seeds, k, mega_draws, seg, rand_seed=0):
    """sample_neighbor with the chains of rows above `mega_draws` draws per lane cut into independent `seg`-draw segments
    positioned by offset skip-ahead (qo_sample_neighbor_gpu_split): must equal sample_neighbor exactly."""
    indptr, indices, seeds = _i64(indptr), _i64(indices), _i64(seeds)
    counts, out_ptr, tot = sample_counts(indptr, seeds, k)
    out = np.zeros(tot, np.int64)
    _lib.qo_sample_neighbor_gpu_split(int(rand_seed), int(k), seeds.shape[0], _p(seeds), _p(indptr), _p(indices),
                                      _p(out_ptr), _p(out), int(mega_draws), int(seg))
    return out, counts


def reindex(inputs, outputs, counts):
    """Quiver.reindex_single: (frontier, row_idx, col_idx)."""
    inputs, outputs, counts = _i64(inputs), _i64(outputs), _i64(counts)
    S, tot = inputs.shape[0], outputs.shape[0]
    frontier, row_idx, col_idx = np.zeros(S + tot, np.int64), np.zeros(tot, np.int64), np.zeros(tot, np.int64)
    F = _lib.qo_reindex(_p(inputs), S, _p(outputs), tot, _p(counts), _p(frontier), _p(row_idx), _p(col_idx))
    return frontier[:F].copy(), row_idx, col_idx


def khop(indptr, indices, seeds, sizes, rand_seed=0, with_eid=False):
    """GraphSageSampler.sample: (n_id, batch_size, [(edge_index[2,E], (n_src, n_dst))] outermost hop first).
    with_eid: every hop's tuple gains a third element, the CSR position of each sampled edge."""
    nodes = _i64(seeds)
    adjs = []
    for size in sizes:
        out, cnt, pos = sample_neighbor_pos(indptr, indices, nodes, size, rand_seed)
        frontier, row_idx, col_idx = reindex(nodes, out, cnt)
        hop = (np.stack([col_idx, row_idx]), (frontier.shape[0], nodes.shape[0]))
        adjs.append(hop + (pos, ) if with_eid else hop)
        nodes = frontier
    return nodes, len(seeds), adjs[::-1]


def gather(shards, indices, feature_order=None, pitches=None):
    """rows of the row-concatenation of `shards` (list of 2-D numpy arrays, same row size); invalid ids -> zero rows."""
    indices = _i64(indices)
    shards = [np.ascontiguousarray(s) for s in shards]
    row_bytes = shards[0].shape[1] * shards[0].itemsize if shards[0].ndim == 2 else shards[0].itemsize
    offsets = np.zeros(len(shards) + 1, np.int64)
    for i, s in enumerate(shards):
        offsets[i + 1] = offsets[i] + s.shape[0]
    ptrs = (ctypes.c_void_p * len(shards))(*[s.ctypes.data for s in shards])
    pitch = _i64(pitches if pitches is not None else [row_bytes] * len(shards))
    out = np.zeros((indices.shape[0],) + shards[0].shape[1:], dtype=shards[0].dtype)
    fo = _i64(feature_order) if feature_order is not None else None
    _lib.qo_gather(ptrs, _p(pitch), _p(offsets), len(shards), _p(indices), _p(fo) if fo is not None else None,
                   indices.shape[0], row_bytes, _p(out))
    return out


def cal_next(last_prob, k, indptr, indices):
    last_prob = np.ascontiguousarray(last_prob, dtype=np.float32)
    indptr, indices = _i64(indptr), _i64(indices)
    cur = np.zeros_like(last_prob)
    _lib.qo_cal_next(_p(last_prob), _p(cur), last_prob.shape[0], int(k), _p(indptr), _p(indices))
    return cur


def validate_sample(indptr, indices, seeds, k, counts, out):
    """0 when (counts, out) is a structurally valid sample (tests/cpp/test_quiver_cpu.cpp:32-51), else 1+bad seed."""
    indptr, indices, seeds, counts, out = map(_i64, (indptr, indices, seeds, counts, out))
    return int(_lib.qo_validate_sample(_p(indptr), _p(indices), _p(seeds), seeds.shape[0], int(k), _p(counts), _p(out),
                                       out.shape[0]))


# ---- the reference's own CPU extension (oracle/_ref, built by build_ref.py) -------------------------------------------
_REF_NAMES = ("torch_quiver_ref", "torch_quiver_ref_omp", "torch_quiver_ref_cuda")


def load_reference(openmp=False):
    """Import the reference CPU extension compiled from /root/reference (None if it was never built).

    All builds of the reference register the same pybind11 types, so one process can hold only ONE of them: if one is
    already imported it is returned (each carries the CPU classes; `cpu_quiver_from_csr_array` behaves the same)."""
    for loaded in _REF_NAMES:
        if loaded in sys.modules:
            return sys.modules[loaded]
    name = "torch_quiver_ref_omp" if openmp else "torch_quiver_ref"
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    try:
        import torch  # noqa: F401  (the extension links libtorch)
        return importlib.import_module(name)
    except ImportError:
        return None


def load_reference_cuda():
    """Import the reference's CUDA extension recompiled for sm_100a by build_ref_cuda.py (None if absent or if another
    build of the reference is already imported in this process).  Exposes the reference's `torch_quiver` surface:
    device_quiver_from_csr_array / Quiver.sample_neighbor / reindex_single / ShardTensor / init_p2p, plus the CPU classes."""
    if "torch_quiver_ref_cuda" in sys.modules:
        return sys.modules["torch_quiver_ref_cuda"]
    if any(n in sys.modules for n in _REF_NAMES):
        return None
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    try:
        import torch  # noqa: F401
        return importlib.import_module("torch_quiver_ref_cuda")
    except ImportError:
        return None


def curand_probe(n, pairs):
    """Run NVIDIA's curand_kernel.h on the host: [{seed, subseq, state, draws}] (None if the probe is not built)."""
    exe = os.path.join(_REF_DIR, "curand_probe")
    if not os.path.exists(exe):
        return None
    import json
    args = [exe, str(n)]
    for s, q in pairs:
        args += [str(s), str(q)]
    return json.loads(subprocess.check_output(args))
