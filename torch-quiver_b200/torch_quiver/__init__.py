"""`torch_quiver` -- drop-in mirror of the reference's pybind11 extension module for the sampler + feature-gather hot
path, implemented as a thin adapter over the C ABI of libquiver_b200.so (include/quiver_b200.h).

Mirrored surface (reference: srcs/cpp/src/quiver/torch/module.cpp:16-26 and the register_* functions it calls):

    device_quiver_from_csr_array(indptr, indices, edge_ids, device=0, cuda=False) -> Quiver   quiver_sample.cu:361,503
    Quiver.sample_neighbor(stream_num, vertices, k) -> (neighbors, counts)                     quiver_sample.cu:113,507
    Quiver.reindex_single(inputs, outputs, counts) -> (frontier, row_idx, col_idx)             quiver_sample.cu:305,511
    Quiver.sample_sub(stream_num, vertices, k) -> (frontier, row_idx, col_idx)                 quiver_sample.cu:257,505
    Quiver.cal_neighbor_prob(stream_num, last_prob, cur_prob, k)                               quiver_sample.cu:100,509
    ShardTensor / ShardTensorItem / init_p2p / can_device_access_peer                          quiver_feature.cu:431-473

PyTorch is used for device memory and streams only.  Deliberate deviations from the reference (SURVEY.md 8(b')):
work is enqueued on torch's CURRENT stream (the reference sampler uses a private pool, quiver_sample.cu:116-117);
CUDA errors raise RuntimeError instead of exit(1); invalid gather ids give zero rows instead of uninitialised memory;
duplicate seeds are merged by reindex (the reference GPU path does too, its CPU path does not); shards are freed.
CPU sampling (cpu_quiver_from_csr_array / CPUQuiver) is not part of this build: there is no CPU path at all.
"""
import ctypes
import weakref
from ctypes import byref, c_int, c_int64, c_void_p

import torch

from . import _lib
from ._lib import QV_MAX_HOPS, QV_MAX_SHARDS, QuiverError, ShardTable, Unsupported, check, lib

__all__ = [
    "device_quiver_from_csr_array", "Quiver", "ShardTensor", "ShardTensorItem", "init_p2p", "can_device_access_peer",
    "cpu_quiver_from_csr_array", "cpu_quiver_from_edge_index", "QuiverError",
]


def _stream(device):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _load_compiled_calls():
    """The compiled call path of the per-step entry points (csrc/pybind/torch_quiver_pybind.cpp: khop_raw), if it is built.

    Same C objects, same C calls as the ctypes code below -- only the host work around them (allocation, pointer tables,
    argument marshalling, result tuples) runs compiled, which end to end is worth ~15 us of GPU idle time per step.
    QUIVER_B200_COMPILED_CALLS=0 keeps everything on ctypes (A-B switch; a missing build does the same)."""
    import importlib.util
    import os
    import sysconfig
    if os.environ.get("QUIVER_B200_COMPILED_CALLS", "1") == "0":
        return None
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "torch_quiver_pybind",
                        "torch_quiver_pb" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        return None
    try:
        spec = importlib.util.spec_from_file_location("torch_quiver_pb", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod if mod.abi_version() == lib.qv_abi_version() and hasattr(mod, "khop_raw") else None
    except Exception:  # a stale or foreign build: the ctypes path serves every call
        return None


_compiled = _load_compiled_calls()


def _ptr(t):
    return c_void_p(t.data_ptr())


def _check_long_cuda(t, name, device=None):
    if not isinstance(t, torch.Tensor) or t.dtype != torch.int64:
        raise RuntimeError(f"{name} must be a torch.long tensor")  # reference: data_ptr<int64_t>() throws
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if device is not None and t.device.index != device:
        raise RuntimeError(f"{name} is on cuda:{t.device.index} but this object lives on cuda:{device}")
    return t.contiguous()


# ----------------------------------------------------------------------------------------------------------------------
# peer access
# ----------------------------------------------------------------------------------------------------------------------
def can_device_access_peer(src_device_index, dst_device_index):
    """torch_quiver.can_device_access_peer -- quiver_feature.cu:422-428."""
    ok = c_int(0)
    check(lib.qv_can_device_access_peer(int(src_device_index), int(dst_device_index), byref(ok)))
    return bool(ok.value)


def init_p2p(devices):
    """torch_quiver.init_p2p -- quiver_feature.cu:378-421 (idempotent here)."""
    devices = [int(d) for d in devices]
    arr = (c_int * max(len(devices), 1))(*devices)
    n = c_int(0)
    check(lib.qv_init_p2p(arr, len(devices), byref(n)))
    return n.value


# ----------------------------------------------------------------------------------------------------------------------
# sampler
# ----------------------------------------------------------------------------------------------------------------------
class Quiver:
    """Device CSR + sampler scratch.  Mirrors `torch_quiver.Quiver` (class TorchQuiver, quiver_sample.cu:77-357)."""

    def __init__(self, indptr, indices, device, cuda, edge_ids=None):
        self.device = int(device)
        self.rand_seed = 0  # the reference hard-codes 0 (quiver.cu.hpp:392); change it to draw a different sample
        self._keep = []  # tensors whose memory the C object borrows
        self._khop_plans = {}  # (S, sizes) -> cached bounds / ctypes scratch of sample_khop
        self._registered = None
        self._handle = c_void_p()
        if indptr.dim() != 1 or indices.dim() != 1:
            raise RuntimeError("check_eq failed")  # reference: check_eq(dim, 1), quiver_sample.cu:373-376
        if indptr.dtype != torch.int64 or indices.dtype != torch.int64:
            raise RuntimeError("indptr / indices must be torch.long")
        if indptr.numel() < 1:
            raise RuntimeError("indptr must hold at least one entry")
        dev = torch.device("cuda", self.device)
        # indptr always lives in HBM (quiver_sample.cu:401-407)
        indptr_d = indptr.to(dev).contiguous()
        self._keep.append(indptr_d)
        if cuda or indices.is_cuda:
            indices_d = indices.to(dev).contiguous()
            self._keep.append(indices_d)
            indices_ptr = indices_d.data_ptr()
        else:
            # UVA / zero-copy: alias the caller's CPU tensor (quiver_sample.cu:413-421); the caller keeps it alive
            indices_c = indices.contiguous()
            self._keep.append(indices_c)
            alias = c_void_p()
            if indices_c.numel() > 0:
                check(lib.qv_host_register(self.device, _ptr(indices_c), indices_c.numel() * 8, byref(alias)))
                self._registered = indices_c
            indices_ptr = alias.value or 0
        self.node_count = indptr.numel() - 1
        self.edge_count = indices.numel()
        check(lib.qv_sampler_create(self.device, _ptr(indptr_d), self.node_count, c_void_p(indices_ptr),
                                    self.edge_count, byref(self._handle)))
        registered = [self._registered.data_ptr()] if self._registered is not None else []
        # edge ids: used iff there is one per edge (quiver_sample.cu:385-387 `use_eid`); HBM copy or zero-copy alias, like
        # `indices` (quiver_sample.cu:434-453).  They feed the opt-in e_id outputs; without them e_id = CSR position.
        self.has_edge_ids = False
        if edge_ids is not None and edge_ids.dim() == 1 and edge_ids.numel() == self.edge_count and self.edge_count > 0:
            if edge_ids.dtype != torch.int64:
                raise RuntimeError("edge_ids must be torch.long")
            if cuda or edge_ids.is_cuda:
                eid_d = edge_ids.to(dev).contiguous()
                self._keep.append(eid_d)
                eid_ptr = eid_d.data_ptr()
            else:
                eid_c = edge_ids.contiguous()
                self._keep.append(eid_c)
                alias = c_void_p()
                check(lib.qv_host_register(self.device, _ptr(eid_c), eid_c.numel() * 8, byref(alias)))
                registered.append(eid_c.data_ptr())
                eid_ptr = alias.value
            check(lib.qv_sampler_set_edge_ids(self._handle, c_void_p(eid_ptr)))
            self.has_edge_ids = True
        self._finalizer = weakref.finalize(self, Quiver._destroy, self._handle.value, tuple(registered))

    @staticmethod
    def _destroy(handle, registered_ptrs):
        if handle:
            lib.qv_sampler_destroy(c_void_p(handle))
        for ptr in registered_ptrs:
            lib.qv_host_unregister(c_void_p(ptr))

    def set_fast(self, enabled=True):
        """Extension: O(k)-per-row sampling that is NOT the reference's random stream (see qv_sampler_set_fast)."""
        check(lib.qv_sampler_set_fast(self._handle, 1 if enabled else 0))
        self.fast = bool(enabled)

    # -- Quiver.sample_neighbor(stream_num, vertices, k) ------------------------------------------------------------
    def sample_neighbor(self, stream_num, vertices, k, return_eid=False):
        """(neighbors, counts); with return_eid=True (extension) also the edge id of every sampled
        neighbour: its CSR position, or edge_ids[position] when the object was built with edge ids."""
        v = _check_long_cuda(vertices, "vertices", self.device)
        S = v.numel()
        counts = torch.empty(S, dtype=torch.int64, device=v.device)
        out_ptr = torch.empty(S, dtype=torch.int64, device=v.device)
        total = c_int64(0)
        st = _stream(self.device)
        check(lib.qv_sample_count(self._handle, _ptr(v), S, int(k), _ptr(counts), _ptr(out_ptr), byref(total), st))
        neighbors = torch.empty(total.value, dtype=torch.int64, device=v.device)
        eid = torch.empty(total.value, dtype=torch.int64, device=v.device) if return_eid else None
        check(lib.qv_sample_fill(self._handle, _ptr(v), S, int(k), int(self.rand_seed), _ptr(out_ptr), _ptr(neighbors),
                                 _ptr(eid) if return_eid else c_void_p(0), st))
        if return_eid:
            return neighbors, counts, eid
        return neighbors, counts

    # -- Quiver.reindex_single(inputs, outputs, counts) -------------------------------------------------------------
    def reindex_single(self, inputs, outputs, counts):
        i = _check_long_cuda(inputs, "inputs", self.device)
        o = _check_long_cuda(outputs, "outputs", self.device)
        c = _check_long_cuda(counts, "counts", self.device)
        S, tot = i.numel(), o.numel()
        if c.numel() != S:
            raise RuntimeError("counts must have one entry per input")
        frontier = torch.empty(S + tot, dtype=torch.int64, device=i.device)
        row_idx = torch.empty(tot, dtype=torch.int64, device=i.device)
        col_idx = torch.empty(tot, dtype=torch.int64, device=i.device)
        n_frontier = c_int64(0)
        check(lib.qv_reindex(self._handle, _ptr(i), S, _ptr(o), tot, _ptr(c), _ptr(frontier), _ptr(row_idx),
                             _ptr(col_idx), byref(n_frontier), _stream(self.device)))
        return frontier[:n_frontier.value], row_idx, col_idx

    # -- Quiver.sample_sub(stream_num, vertices, k) -----------------------------------------------------------------
    def sample_sub(self, stream_num, vertices, k):
        """sample_neighbor + reindex_single fused (quiver_sample.cu:257-304): ONE C call (a one-hop qv_khop) and one host
        synchronisation; k = -1 (no static bound) takes the two calls."""
        if int(k) >= 0 and vertices.numel() > 0:
            try:
                n_id, hops = self.sample_khop(vertices, [int(k)])
            except Unsupported:
                pass
            else:
                edge_index = hops[0][0]
                return n_id, edge_index[1], edge_index[0]  # (frontier, row_idx, col_idx)
        out, cnt = self.sample_neighbor(stream_num, vertices, k)
        return self.reindex_single(vertices, out, cnt)

    # -- Quiver.cal_neighbor_prob(stream_num, last_prob, cur_prob, k) -----------------------------------------------
    def cal_neighbor_prob(self, stream_num, last_prob, cur_prob, k):
        for t in (last_prob, cur_prob):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("cal_neighbor_prob expects contiguous float32 CUDA tensors")
        check(lib.qv_cal_neighbor_prob(self._handle, _ptr(last_prob), _ptr(cur_prob), cur_prob.numel(), int(k),
                                       _stream(self.device)))

    # -- fused k-hop (ours): every hop enqueued back to back, one host synchronisation --------------------------------
    def sample_khop(self, seeds, sizes, gather=None, with_eid=False):
        """All hops of GraphSageSampler.sample (sage_sampler.py:118-147) in one C call.

        Returns (n_id, [(edge_index[2, E_l], n_src_l, n_dst_l) for l in hops, innermost first]).
        Raises `Unsupported` when a size is negative or the static bound is too large (callers fall back to the
        per-hop calls).

        gather=(shard_tensor, feature_order or None): also gather the feature rows of n_id behind the last hop, without
        a host round trip in between (qv_khop_gather); returns (n_id, hops, x) with x = shard_tensor[feature_order[n_id]]
        produced stream-ordered on the current stream.  x is a VIEW of a buffer allocated before the frontier size is
        known: min(S * prod(1 + size), node_count + S) rows (a frontier holds distinct nodes), at most
        QUIVER_B200_FUSED_GATHER_MAX bytes (default 2 GiB, beyond that `Unsupported` -> callers take the two exact
        calls).  The whole block stays alive while the caller holds x; `x = x.clone()` releases it.

        with_eid=True: every hop tuple gains a 4th element, the e_id of its edges (CSR position or user edge id)."""
        # seeds: a CUDA tensor, or a PINNED host tensor -- pinned memory is device-visible at the same address under unified
        # addressing, so the kernels of hop 0 read the seeds straight over PCIe (8 KB) and the call needs no staging copy
        if _compiled is not None and isinstance(seeds, torch.Tensor) and seeds.dtype == torch.int64 and seeds.is_contiguous():
            if gather is None:
                res = _compiled.khop_raw(self._handle.value, seeds, sizes, int(self.rand_seed), self.device, self.node_count, 0,
                                         None, 0, (), torch.int64, 0, with_eid, 0)
                if res is not None:
                    return res[0], res[1]
            else:
                store, feature_order = gather
                table, dtype, row_shape, row_bytes = store._gather_plan(self.device)
                res = _compiled.khop_raw(self._handle.value, seeds, sizes, int(self.rand_seed), self.device, self.node_count,
                                         ctypes.addressof(table), feature_order, row_bytes, row_shape, dtype,
                                         int(store.gather_variant), with_eid, _FUSED_GATHER_MAX_BYTES)
                if res is not None:
                    return res
            # None: the request is one this path does not serve -- the code below raises the matching error
        if isinstance(seeds, torch.Tensor) and not seeds.is_cuda and seeds.dtype == torch.int64 and seeds.is_pinned() \
                and seeds.is_contiguous():
            v = seeds
        else:
            v = _check_long_cuda(seeds, "seeds", self.device)
        out_dev = torch.device("cuda", self.device)
        n_hops = len(sizes)
        S = v.numel()
        key = (S, tuple(int(x) for x in sizes))
        plan = self._khop_plans.get(key)
        if plan is None:
            if not 1 <= n_hops <= QV_MAX_HOPS:
                raise Unsupported(_lib.QV_ERR_UNSUPPORTED, f"{n_hops} hops")
            sz = (c_int64 * n_hops)(*key[1])
            bn = (c_int64 * (n_hops + 1))()
            be = (c_int64 * n_hops)()
            check(lib.qv_khop_bounds(S, sz, n_hops, bn, be))
            # one allocation per call: [n_id | edge_buf[0] | edge_buf[1] | ...], every piece 16-byte aligned
            offs, total = [], (max(bn[n_hops], 1) + 1) // 2 * 2
            for h in range(n_hops):
                offs.append(total)
                total += max(2 * be[h], 2)
            eoffs = []  # e_id regions (only allocated when asked for): appended behind the edge buffers
            etotal = total
            for h in range(n_hops):
                eoffs.append(etotal)
                etotal += max(be[h], 2)
            plan = (sz, bn[n_hops], offs, total, (c_void_p * n_hops)(), (c_int64 * (n_hops + 1))(), (c_int64 * n_hops)(),
                    eoffs, etotal, (c_void_p * n_hops)())
            if len(self._khop_plans) > 64:
                self._khop_plans.clear()
            self._khop_plans[key] = plan
        sz, n_id_cap, offs, total, buf_ptrs, out_nodes, out_edges, eoffs, etotal, eid_ptrs = plan
        arena = torch.empty(etotal if with_eid else total, dtype=torch.int64, device=out_dev)
        base = arena.data_ptr()
        for h in range(n_hops):
            buf_ptrs[h] = base + 8 * offs[h]
            if with_eid:
                eid_ptrs[h] = base + 8 * eoffs[h]
        eid_arg = eid_ptrs if with_eid else None
        x = None
        if gather is None:
            check(lib.qv_khop(self._handle, _ptr(v), S, sz, n_hops, int(self.rand_seed), c_void_p(base), buf_ptrs,
                              eid_arg, out_nodes, out_edges, _stream(self.device)))
        else:
            store, feature_order = gather
            if torch.cuda.current_device() != self.device:
                raise RuntimeError("sample_khop(gather=...) must run with the sampler's device current")
            table, dtype, row_shape, row_bytes = store._gather_plan(self.device)
            x_rows = max(min(n_id_cap, self.node_count + S), 1)
            if x_rows * row_bytes > _FUSED_GATHER_MAX_BYTES:
                # the output is sized for the STATIC frontier bound (the real size is not known when the gather is
                # enqueued); beyond this the two separate calls, which allocate exactly, are the better trade
                raise Unsupported(_lib.QV_ERR_UNSUPPORTED, "fused gather buffer would exceed QUIVER_B200_FUSED_GATHER_MAX")
            order_ptr = c_void_p(0)
            if feature_order is not None:
                order_ptr = _ptr(_check_long_cuda(feature_order, "feature_order", self.device))
            x = torch.empty([x_rows] + row_shape, dtype=dtype, device=out_dev)
            check(lib.qv_khop_gather(self._handle, _ptr(v), S, sz, n_hops, int(self.rand_seed), c_void_p(base), buf_ptrs,
                                     eid_arg, byref(table), order_ptr, row_bytes, _ptr(x), x_rows,
                                     int(store.gather_variant), out_nodes, out_edges, _stream(self.device)))
        hops = []
        for h in range(n_hops):
            E = out_edges[h]
            hop = (arena[offs[h]:offs[h] + 2 * E].view(2, E), out_nodes[h + 1], out_nodes[h])
            hops.append(hop + (arena[eoffs[h]:eoffs[h] + E], ) if with_eid else hop)
        if gather is None:
            return arena[:out_nodes[n_hops]], hops
        return arena[:out_nodes[n_hops]], hops, x[:out_nodes[n_hops]]


def device_quiver_from_csr_array(indptr, indices, edge_ids=None, device=0, cuda=False):
    """torch_quiver.device_quiver_from_csr_array -- quiver_sample.cu:361-461.  `edge_ids` is used iff it holds one id per
    edge (quiver_sample.cu:385-387); the reference then plumbs it and returns an empty e_id anyway
    (sage_sampler.py:143) -- here it feeds the opt-in e_id outputs (sample_neighbor(return_eid=True),
    sample_khop(with_eid=True))."""
    return Quiver(indptr, indices, device, cuda, edge_ids)


def cpu_quiver_from_csr_array(*_args, **_kwargs):
    raise NotImplementedError("the B200 build has no CPU sampling path (reference: srcs/cpp/src/quiver/quiver.cpp); "
                              "use mode='GPU' or mode='UVA'")


cpu_quiver_from_edge_index = cpu_quiver_from_csr_array


# ----------------------------------------------------------------------------------------------------------------------
# feature shards
# ----------------------------------------------------------------------------------------------------------------------
_ELEMENT_DTYPE = {1: torch.uint8, 2: torch.float16, 4: torch.float32, 8: torch.float64}  # quiver_feature.cu:262-267


import os as _os

_FUSED_GATHER_MAX_BYTES = int(_os.environ.get("QUIVER_B200_FUSED_GATHER_MAX", str(2 << 30)))
_PITCH_ALIGN = int(_os.environ.get("QUIVER_B200_PITCH_ALIGN", "16"))  # bytes; 64 aligns rows to DRAM access granules


def _alloc_bytes(nbytes):
    """qv_malloc rounds every block of >= 2 MiB up to a 2 MiB multiple itself (peer page size, see quiver_b200.h)."""
    return max(int(nbytes), 16)


def _pitch_for(row_bytes):
    return (row_bytes + _PITCH_ALIGN - 1) // _PITCH_ALIGN * _PITCH_ALIGN


class _RawDeviceMemory:
    """__cuda_array_interface__ carrier: lets torch view memory the library owns (no copy, no ownership transfer)."""

    def __init__(self, ptr, n_elems, typestr):
        self.__cuda_array_interface__ = {"shape": (n_elems, ), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


_TYPESTR = {torch.float32: "<f4", torch.float16: "<f2", torch.float64: "<f8", torch.uint8: "|u1", torch.int32: "<i4",
            torch.int64: "<i8", torch.int16: "<i2", torch.int8: "|i1"}


def _device_view(ptr, rows, pitch, row_bytes, dtype, device, row_shape):
    if rows == 0:
        return torch.empty([0] + row_shape, dtype=dtype, device=f"cuda:{device}")
    esz = torch.empty(0, dtype=dtype).element_size()
    carrier_dtype = dtype if dtype in _TYPESTR else {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[esz]
    with torch.cuda.device(device):
        flat = torch.as_tensor(_RawDeviceMemory(ptr, rows * pitch // esz, _TYPESTR[carrier_dtype]), device=f"cuda:{device}")
    if carrier_dtype != dtype:
        flat = flat.view(dtype)
    return flat.view(rows, pitch // esz)[:, :row_bytes // esz].unflatten(1, row_shape) if len(row_shape) != 1 else \
        flat.view(rows, pitch // esz)[:, :row_bytes // esz]


class ShardTensorItem:
    """CUDA-IPC carrier of one GPU shard -- class ShardTensorItem, quiver_feature.cu:20-55."""

    def __init__(self):
        self.device = -1
        self.element_size = 4
        self.mem_handle = b"\0" * _lib.QV_IPC_HANDLE_BYTES
        self.shape = []

    def share_ipc(self):
        return self.device, self.element_size, self.mem_handle, list(self.shape)

    def from_ipc(self, ipc_data):
        self.device, self.element_size, handle, shape = ipc_data
        self.mem_handle = bytes(handle)
        self.shape = list(shape)


class _Shard:
    __slots__ = ("device", "ptr", "rows", "pitch", "owned", "ipc_opened", "host_tensor", "host_base", "shape",
                 "open_device")

    def __init__(self, device, ptr, rows, pitch, owned=False, ipc_opened=False, host_tensor=None, shape=None,
                 open_device=None):
        self.device, self.ptr, self.rows, self.pitch = device, ptr, rows, pitch
        self.owned, self.ipc_opened, self.host_tensor, self.shape = owned, ipc_opened, host_tensor, shape
        self.open_device = open_device  # the device in whose context an IPC handle was opened (and must be closed)
        self.host_base = host_tensor.data_ptr() if (host_tensor is not None and ptr) else 0  # what we registered


class ShardTensor:
    """Row-sharded feature table readable from one GPU -- class ShardTensor, quiver_feature.cu:57-376.

    Shards are appended in row order; shard s holds rows [offset[s], offset[s+1]).  device >= 0: rows are copied into
    that GPU's HBM (padded to a 16-byte pitch) and read in the gather kernel through a peer-mapped pointer;
    device == -1: the caller's CPU tensor is registered and read zero-copy over PCIe (it is aliased, not copied --
    keep it alive, as with the reference)."""

    def __init__(self, device):
        self.device_ = int(device)
        self.shards = []
        self.offset_list_ = [0]
        self.shape_ = []
        self.element_size = 4
        self.dtype = None
        self.gather_variant = 0
        self._finalizer = weakref.finalize(self, ShardTensor._release, self.shards)

    @staticmethod
    def _release(shards):
        for sh in shards:
            try:
                if sh.owned and sh.ptr:
                    lib.qv_free(sh.device, c_void_p(sh.ptr))
                elif sh.ipc_opened and sh.ptr:
                    lib.qv_ipc_close_handle(sh.device if sh.open_device is None else sh.open_device, c_void_p(sh.ptr))
                elif sh.host_base:
                    # a registration that outlives its memory poisons later cudaMemcpy calls on reused addresses
                    lib.qv_host_unregister(c_void_p(sh.host_base))
            except Exception:  # interpreter shutdown
                pass
        shards.clear()

    # -- bookkeeping shared by both append flavours -----------------------------------------------------------------
    def _admit(self, shape, element_size, dtype):
        shape = [int(x) for x in shape]
        if len(shape) < 1:
            raise RuntimeError("a shard needs at least one dimension")
        if not self.shape_:
            self.shape_ = [0] + shape[1:]
            self.element_size = int(element_size)
            self.dtype = dtype
        elif shape[1:] != self.shape_[1:] or int(element_size) != self.element_size:
            raise RuntimeError(f"shard shape {shape} / element size {element_size} does not match "
                               f"{self.shape_} / {self.element_size}")
        if len(self.shards) >= QV_MAX_SHARDS:
            raise RuntimeError(f"at most {QV_MAX_SHARDS} shards per ShardTensor")
        self.shape_[0] += shape[0]
        self.offset_list_.append(self.offset_list_[-1] + shape[0])

    def _row_bytes(self):
        return self.stride(0) * self.element_size

    def append(self, item, target_device=None):
        if isinstance(item, ShardTensorItem):
            return self._append_item(item)
        return self._append_tensor(item, int(target_device))

    def _append_tensor(self, tensor, target_device):
        if tensor.is_cuda:
            # extension: the reference refuses (CHECK_CPU, quiver_feature.cu:19,147), which caps a table at host-memory
            # size; rows already in HBM are copied device-to-device into the shard
            return self._append_device_tensor(tensor, target_device)
        tensor = tensor if tensor.is_contiguous() else tensor.contiguous()
        self._admit(tensor.shape, tensor.element_size(), tensor.dtype)
        rows, row_bytes = tensor.shape[0], self._row_bytes()
        if target_device >= 0:
            pitch = _pitch_for(row_bytes)
            ptr = c_void_p()
            check(lib.qv_malloc(target_device, _alloc_bytes(rows * pitch), byref(ptr)))
            check(lib.qv_upload_rows(target_device, ptr, pitch, _ptr(tensor), row_bytes, row_bytes, rows))
            if target_device != self.device_ and can_device_access_peer(self.device_, target_device):
                init_p2p([self.device_, target_device])
            self.shards.append(_Shard(target_device, ptr.value, rows, pitch, owned=True, shape=list(tensor.shape)))
        else:
            alias = c_void_p()
            if tensor.numel() > 0:
                check(lib.qv_host_register(self.device_, _ptr(tensor), tensor.numel() * tensor.element_size(),
                                           byref(alias)))
            self.shards.append(_Shard(-1, alias.value or 0, rows, row_bytes, host_tensor=tensor,
                                      shape=list(tensor.shape)))

    def _append_device_tensor(self, tensor, target_device):
        if target_device < 0:
            raise RuntimeError("a CUDA tensor cannot become the pinned-host tier")
        tensor = tensor if tensor.is_contiguous() else tensor.contiguous()
        view = self.append_empty(tensor.shape[0], list(tensor.shape[1:]), tensor.dtype, target_device)
        row_bytes = self._row_bytes()
        pitch = self.shards[-1].pitch
        src_dev = tensor.device.index
        with torch.cuda.device(src_dev):
            st = _stream(src_dev)
            check(lib.qv_copy_rows_device(src_dev, c_void_p(self.shards[-1].ptr), pitch, _ptr(tensor), row_bytes, row_bytes,
                                          tensor.shape[0], st))
            torch.cuda.current_stream(src_dev).synchronize()
        del view

    def append_empty(self, rows, row_shape, dtype, target_device):
        """Extension: create the next shard IN PLACE in `target_device`'s HBM and return a torch view of it
        ([rows, *row_shape], row stride = the shard's 16-byte-aligned pitch) for the caller to fill on the device -- how
        a table larger than host memory (mag240m: 750 GB over 8 GPUs) is built.  The shard owns the memory
        (cudaMalloc: exportable over CUDA IPC); the view must not outlive the ShardTensor."""
        rows = int(rows)
        dtype_size = torch.empty(0, dtype=dtype).element_size()
        self._admit([rows] + [int(d) for d in row_shape], dtype_size, dtype)
        row_bytes = self._row_bytes()
        pitch = _pitch_for(row_bytes)
        if pitch % dtype_size:
            raise RuntimeError("row pitch is not a multiple of the element size")
        ptr = c_void_p()
        check(lib.qv_malloc(target_device, _alloc_bytes(rows * pitch), byref(ptr)))
        if target_device != self.device_ and can_device_access_peer(self.device_, target_device):
            init_p2p([self.device_, target_device])
        self.shards.append(_Shard(target_device, ptr.value, rows, pitch, owned=True, shape=[rows] + list(row_shape)))
        return _device_view(ptr.value, rows, pitch, row_bytes, dtype, target_device, list(row_shape))

    def _append_item(self, item):
        self._admit(item.shape, item.element_size, _ELEMENT_DTYPE.get(int(item.element_size)))
        row_bytes = self._row_bytes()
        ptr = c_void_p()
        handle = (ctypes.c_ubyte * _lib.QV_IPC_HANDLE_BYTES).from_buffer_copy(bytes(item.mem_handle))
        # open in the context of the device that will dereference it (quiver_feature.cu:122-134)
        check(lib.qv_ipc_open_handle(self.device_, handle, byref(ptr)))
        self.shards.append(_Shard(int(item.device), ptr.value, int(item.shape[0]), _pitch_for(row_bytes),
                                  ipc_opened=True, shape=list(item.shape), open_device=self.device_))

    def adopt(self, other):
        """Move the (single) shard of another ShardTensor to the end of this one, keeping ownership of its memory
        (used when ranks assemble a table in rank order: quiver.shard_tensor.build_from_ranks)."""
        if len(other.shards) != 1:
            raise RuntimeError("adopt() takes a ShardTensor holding exactly one shard")
        sh = other.shards[0]
        self._admit(sh.shape, other.element_size, other.dtype)
        self.shards.append(sh)
        other.shards.clear()

    # -- ShardTensor.__getitem__(indices) ---------------------------------------------------------------------------
    def _table(self, current_device):
        t = ShardTable()
        t.n_shards = len(self.shards)
        for s, sh in enumerate(self.shards):
            t.row_begin[s] = self.offset_list_[s]
            t.ptr[s] = sh.ptr
            t.pitch[s] = sh.pitch
            t.accessible[s] = 1 if (sh.device < 0 or sh.device == current_device
                                    or can_device_access_peer(current_device, sh.device)) else 0
        t.row_begin[len(self.shards)] = self.offset_list_[-1]
        return t

    def _gather_plan(self, current):
        """(shard table as seen from device `current`, dtype, row shape, row bytes) -- cached per device."""
        if not self.shards:
            raise RuntimeError("ShardTensor is empty")
        dtype = self.dtype or _ELEMENT_DTYPE.get(self.element_size)
        if dtype is None:
            raise RuntimeError(f"unsupported element size {self.element_size}")
        cache = getattr(self, "_table_cache", None)
        if cache is None or cache[0] != (current, len(self.shards)):
            self._table_cache = ((current, len(self.shards)), self._table(current))
        return self._table_cache[1], dtype, list(self.shape_[1:]), self._row_bytes()

    def gather(self, indices, feature_order=None, out=None):
        """`self[indices]` with the optional `feature_order[idx]` indirection folded into the kernel."""
        if _compiled is not None and out is None and isinstance(indices, torch.Tensor) and indices.is_cuda \
                and indices.is_contiguous():
            table, dtype, row_shape, row_bytes = self._gather_plan(indices.device.index)
            return _compiled.gather_raw(ctypes.addressof(table), indices, feature_order, row_bytes, row_shape, dtype,
                                        int(self.gather_variant))
        idx = _check_long_cuda(indices, "indices")
        current = idx.device.index
        n = idx.numel()
        table, dtype, row_shape, _ = self._gather_plan(current)
        if out is None:
            out = torch.empty([n] + row_shape, dtype=dtype, device=idx.device)
        order_ptr = c_void_p(0)
        if feature_order is not None:
            fo = _check_long_cuda(feature_order, "feature_order", current)
            order_ptr = _ptr(fo)
        if torch.cuda.current_device() == current:
            check(lib.qv_gather(byref(table), _ptr(idx), order_ptr, n, self._row_bytes(), _ptr(out),
                                int(self.gather_variant), _stream(current)))
        else:
            with torch.cuda.device(current):
                check(lib.qv_gather(byref(table), _ptr(idx), order_ptr, n, self._row_bytes(), _ptr(out),
                                    int(self.gather_variant), _stream(current)))
        return out

    def __getitem__(self, indices):
        return self.gather(indices)

    # -- introspection (quiver_feature.cu:304-333, 352) -------------------------------------------------------------
    def shape(self):
        return list(self.shape_)

    def device(self):
        return self.device_

    def size(self, dim):
        return self.shape_[dim] if self.shape_ else 0

    def stride(self, dim):
        res = 1
        for d in self.shape_[dim + 1:]:
            res *= d
        return res

    def numel(self):
        res = 1
        for d in self.shape_:
            res *= d
        return res

    def device_count(self):
        return len(self.shards)

    def share_ipc(self):
        """One ShardTensorItem per GPU shard (quiver_feature.cu:335-350); the host tier travels separately."""
        items = []
        for sh in self.shards:
            if sh.device < 0:
                continue
            if not sh.owned:
                raise RuntimeError("only the process that created a GPU shard can export it")
            handle = (ctypes.c_ubyte * _lib.QV_IPC_HANDLE_BYTES)()
            check(lib.qv_ipc_get_handle(sh.device, c_void_p(sh.ptr), handle))
            item = ShardTensorItem()
            item.device, item.element_size, item.mem_handle, item.shape = sh.device, self.element_size, bytes(handle), \
                list(sh.shape)
            items.append(item)
        return items

    def unregister(self, cpu_tensor):
        """ShardTensor.unregister(cpu_tensor) -- quiver_feature.cu:354-360."""
        check(lib.qv_host_unregister(_ptr(cpu_tensor)))
        for sh in self.shards:
            if sh.host_base == cpu_tensor.data_ptr():
                sh.host_base = 0

    def move_host_tier_to_shared_memory(self):
        """`tensor.share_memory_()` on the zero-copy host tier WITHOUT leaving a dangling registration: the storage
        moves to a new mapping, so the old range is unregistered first and the new one registered after.  (The
        reference calls share_memory_() on a registered tensor, feature.py:383-384, which leaves the parent process
        reading freed memory.)"""
        for sh in self.shards:
            if sh.host_tensor is None or sh.host_tensor.is_shared() or sh.host_tensor.numel() == 0:
                continue
            if sh.host_base:
                check(lib.qv_host_unregister(c_void_p(sh.host_base)))
                sh.host_base, sh.ptr = 0, 0
            sh.host_tensor.share_memory_()
            alias = c_void_p()
            check(lib.qv_host_register(self.device_, _ptr(sh.host_tensor),
                                       sh.host_tensor.numel() * sh.host_tensor.element_size(), byref(alias)))
            sh.ptr, sh.host_base = alias.value, sh.host_tensor.data_ptr()
        self._table_cache = None
