"""ctypes binding of libquiver_b200.so -- one prototype per symbol declared in include/quiver_b200.h.

The library is the product's only compute path.  If it is missing this module raises ImportError (there is no CPU or
PyTorch fallback anywhere in the package); if it is present but CUDA is not, every call raises RuntimeError with the
CUDA error text.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

QV_MAX_SHARDS = 16
QV_MAX_HOPS = 8
QV_IPC_HANDLE_BYTES = 64
QV_OK, QV_ERR_INVALID, QV_ERR_CUDA, QV_ERR_NOMEM, QV_ERR_UNSUPPORTED = 0, 1, 2, 3, 4

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QUIVER_B200_LIB", os.path.join(_HERE, "libquiver_b200.so"))


class ShardTable(ctypes.Structure):
    """struct qv_shard_table (include/quiver_b200.h)."""
    _fields_ = [
        ("n_shards", c_int32),
        ("reserved", c_int32),
        ("row_begin", c_int64 * (QV_MAX_SHARDS + 1)),
        ("ptr", c_void_p * QV_MAX_SHARDS),
        ("pitch", c_int64 * QV_MAX_SHARDS),
        ("accessible", c_int32 * QV_MAX_SHARDS),
    ]


class QuiverError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libquiver_b200 error {code}: {message}")
        self.code = code


class Unsupported(QuiverError):
    pass


# name -> (restype, argtypes).  Kept as data so tests can check it against the header.
PROTOTYPES = {
    "qv_abi_version": (c_int, []),
    "qv_last_error": (c_char_p, []),
    "qv_launch_count": (c_int64, []),
    "qv_device_count": (c_int, [POINTER(c_int)]),
    "qv_can_device_access_peer": (c_int, [c_int, c_int, POINTER(c_int)]),
    "qv_init_p2p": (c_int, [POINTER(c_int), c_int, POINTER(c_int)]),
    "qv_malloc": (c_int, [c_int, c_size_t, POINTER(c_void_p)]),
    "qv_free": (c_int, [c_int, c_void_p]),
    "qv_upload_rows": (c_int, [c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t]),
    "qv_memset": (c_int, [c_int, c_void_p, c_int, c_size_t]),
    "qv_copy_rows_device": (c_int, [c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    "qv_host_register": (c_int, [c_int, c_void_p, c_size_t, POINTER(c_void_p)]),
    "qv_host_unregister": (c_int, [c_void_p]),
    "qv_ipc_get_handle": (c_int, [c_int, c_void_p, c_void_p]),
    "qv_ipc_open_handle": (c_int, [c_int, c_void_p, POINTER(c_void_p)]),
    "qv_ipc_close_handle": (c_int, [c_int, c_void_p]),
    "qv_gather": (c_int, [POINTER(ShardTable), c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "qv_sampler_create": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, POINTER(c_void_p)]),
    "qv_sampler_destroy": (c_int, [c_void_p]),
    "qv_sample_count": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, POINTER(c_int64), c_void_p]),
    "qv_sampler_set_edge_ids": (c_int, [c_void_p, c_void_p]),
    "qv_sample_fill": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "qv_reindex": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                           POINTER(c_int64), c_void_p]),
    "qv_khop_bounds": (c_int, [c_int64, POINTER(c_int64), c_int, POINTER(c_int64), POINTER(c_int64)]),
    "qv_khop": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_int64), c_int, c_uint64, c_void_p, POINTER(c_void_p),
                        POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64), c_void_p]),
    "qv_khop_gather": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_int64), c_int, c_uint64, c_void_p, POINTER(c_void_p),
                               POINTER(c_void_p), POINTER(ShardTable), c_void_p, c_int64, c_void_p, c_int64, c_int,
                               POINTER(c_int64), POINTER(c_int64), c_void_p]),
    "qv_sampler_set_fast": (c_int, [c_void_p, c_int]),
    "qv_cal_neighbor_prob": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"libquiver_b200.so not found at {LIB_PATH}. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or torch-quiver_b200/csrc/build.sh). There is no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def check(rc):
    if rc != QV_OK:
        msg = (lib.qv_last_error() or b"").decode("utf-8", "replace")
        raise (Unsupported if rc == QV_ERR_UNSUPPORTED else QuiverError)(rc, msg)


def launch_count():
    return int(lib.qv_launch_count())
