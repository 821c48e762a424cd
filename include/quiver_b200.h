/*
 * quiver_b200.h -- C ABI of libquiver_b200.so, the B200-native (sm_100a) drop-in for torch-quiver's two
 * data-parallel hot paths: the CSR k-hop neighbour sampler and the tiered / sharded feature gather.
 *
 * The reference has no C ABI: its plugin boundary is the pybind11 module `torch_quiver`
 * (srcs/cpp/src/quiver/torch/module.cpp:16-26).  Every entry point below names the reference binding /
 * function it replaces (file:line relative to the reference tree).  The Python mirror of that pybind surface
 * (torch-quiver_b200/torch_quiver/) is a thin ctypes adapter over exactly these symbols; INTEGRATION.md shows
 * the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / pybind types.  Device pointers are raw CUDA device addresses.
 *   - every function returns QV_OK (0) or an error code; qv_last_error() gives the message (thread-local).
 *     CUDA errors are REPORTED, never exit(1) (the reference does exit: include/quiver/common.hpp:18-26).
 *   - `stream` is a cudaStream_t passed as void*; all device work is enqueued on it, in order.  Functions that
 *     must hand a size back to the host (documented below) synchronise that stream once.
 *   - all ids are int64 (torch.long), as in the reference API.
 *   - one qv_sampler per (process, device, stream); objects are not thread-safe (same as the reference:
 *     SURVEY.md 8(b) "Threading").
 */
#ifndef QUIVER_B200_H
#define QUIVER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QV_ABI_VERSION 2 /* 2: edge-id outputs (qv_sample_fill / qv_khop / qv_khop_gather), qv_copy_rows_device */
#define QV_MAX_SHARDS 16 /* 8 GPU shards + pinned-host tier, with headroom (reference tables hold <= 9) */
#define QV_MAX_HOPS 8
#define QV_IPC_HANDLE_BYTES 64 /* CUDA_IPC_HANDLE_SIZE */

typedef void *qv_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define QV_API __attribute__((visibility("default")))
#else
#define QV_API
#endif

enum qv_status {
    QV_OK = 0,
    QV_ERR_INVALID = 1,     /* bad argument (reference: check_eq -> std::runtime_error, common.hpp:6-16) */
    QV_ERR_CUDA = 2,        /* a CUDA runtime call failed */
    QV_ERR_NOMEM = 3,       /* device / pinned allocation failed */
    QV_ERR_UNSUPPORTED = 4, /* valid request this build does not serve (e.g. unbounded fused k-hop) */
};

QV_API int qv_abi_version(void);
QV_API const char *qv_last_error(void);

/* ---------------------------------------------------------------------------------------------------------
 * Devices, peer access, memory tiers
 * ------------------------------------------------------------------------------------------------------- */

/* cudaGetDeviceCount. */
QV_API int qv_device_count(int *count);

/* torch_quiver.can_device_access_peer(src, dst) -- srcs/cpp/src/quiver/cuda/quiver_feature.cu:422-428.
 * *ok = 1 iff both directions report cudaDeviceCanAccessPeer (src == dst counts as accessible). */
QV_API int qv_can_device_access_peer(int src, int dst, int *ok);

/* torch_quiver.init_p2p(devices) -- quiver_feature.cu:378-421.  Enables peer access for every ordered pair
 * that supports it; "already enabled" is not an error.  *n_enabled (optional) = number of directed links now
 * enabled. */
QV_API int qv_init_p2p(const int *devices, int n_devices, int *n_enabled);

/* Device shard storage.  ShardTensor::append(tensor, device>=0) does cudaMalloc + cudaMemcpy(H2D)
 * (quiver_feature.cu:166-174); these are the same two steps, exposed separately, plus the matching free the
 * reference never performs.  qv_upload_rows copies `rows` rows of `row_bytes` from host memory with source
 * pitch `src_pitch` into device memory with destination pitch `dst_pitch` (cudaMemcpy2D), synchronously. */
/* qv_malloc rounds blocks of >= 2 MiB up to a 2 MiB multiple: peers read a block whose size is not a whole number of
 * large pages up to 12x slower (60-105 vs 745 GB/s for random 1 KiB rows, profiles/r2_peer_alloc_granularity.txt). */
QV_API int qv_malloc(int device, size_t bytes, void **dev_ptr);
QV_API int qv_free(int device, void *dev_ptr);
QV_API int qv_upload_rows(int device, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes,
                   size_t rows);
QV_API int qv_memset(int device, void *dst, int value, size_t bytes);
/* Shard creation from DEVICE memory (extension; the reference's append() only takes CPU tensors, CHECK_CPU at
 * quiver_feature.cu:19,147, so a table larger than host memory cannot be staged): device-to-device pitched row copy,
 * asynchronous on `stream`.  A shard can also be created empty with qv_malloc and filled in place by the caller. */
QV_API int qv_copy_rows_device(int device, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes,
                        size_t rows, qv_stream_t stream);

/* Zero-copy host tier: quiverRegister(cudaHostRegisterMapped) + cudaHostGetDevicePointer
 * (include/quiver/quiver.cu.hpp:16-26, quiver_feature.cu:192-199, quiver_sample.cu:413-421).
 * Registers exactly [host_ptr, host_ptr+bytes) (cudaHostRegister accepts unaligned ranges and pins the pages they
 * touch) and returns the device-visible alias.  Registering an already registered range is not an error: registrations
 * are reference counted per base pointer, a longer range on the same base replaces the shorter one, and a range that
 * someone else pinned (torch pin_memory) is used as is and never unregistered here. */
QV_API int qv_host_register(int device, void *host_ptr, size_t bytes, void **dev_ptr);
/* ShardTensor.unregister(cpu_tensor) -- quiver_feature.cu:354-360. */
QV_API int qv_host_unregister(void *host_ptr);

/* CUDA IPC for passing GPU shards through mp.spawn: ShardTensor::share_ipc (quiver_feature.cu:335-350,
 * cudaIpcGetMemHandle) and ShardTensor::append(ShardTensorItem) (quiver_feature.cu:86-143,
 * cudaIpcOpenMemHandle with cudaIpcMemLazyEnablePeerAccess). */
QV_API int qv_ipc_get_handle(int device, void *dev_ptr, unsigned char handle[QV_IPC_HANDLE_BYTES]);
QV_API int qv_ipc_open_handle(int device, const unsigned char handle[QV_IPC_HANDLE_BYTES], void **dev_ptr);
QV_API int qv_ipc_close_handle(int device, void *dev_ptr);

/* ---------------------------------------------------------------------------------------------------------
 * Feature gather ("collect")
 * ------------------------------------------------------------------------------------------------------- */

/* The shard table of one ShardTensor as seen from the gathering device: dev_ptrs_ / offset_list_ /
 * access_book (quiver_feature.cu:362-366, 208-244).  Shard s holds logical rows
 * [row_begin[s], row_begin[s+1]); ptr[s] is a pointer valid on the gathering device (local HBM, peer HBM
 * mapped over NVLink, or a registered host alias) to that shard's first row; pitch[s] is its row pitch in
 * bytes (>= row_bytes; the B200 build pads pitches to 16 B so rows can be moved with 16-byte / bulk copies).
 * accessible[s] == 0 marks a shard the device cannot dereference (the reference's access_book == 0): its rows
 * are zero-filled here and completed by the host-side fallback (srcs/python/quiver/shard_tensor.py:138-152). */
typedef struct qv_shard_table {
    int32_t n_shards;
    int32_t reserved;
    int64_t row_begin[QV_MAX_SHARDS + 1];
    const void *ptr[QV_MAX_SHARDS];
    int64_t pitch[QV_MAX_SHARDS];
    int32_t accessible[QV_MAX_SHARDS];
} qv_shard_table;

/* ShardTensor.__getitem__(indices) -- quiver_feature.cu:246-302 launching quiver_tensor_gather
 * (include/quiver/shard_tensor.cu.hpp:19-61), with Feature.__getitem__'s `feature_order[idx]` indirection
 * (srcs/python/quiver/feature.py:300-301) folded in when feature_order != NULL.
 *   out[i, 0:row_bytes] = shard(j)[j - row_begin(shard(j))],  j = feature_order ? feature_order[indices[i]]
 *                                                                               : indices[i]
 * `out` is a dense [n, row_bytes] buffer on the current device.  Rows whose index is < 0 or >= total rows (or
 * whose shard is not accessible) are written as zeros -- the reference leaves them uninitialised
 * (shard_tensor.cu.hpp:49).  Pure byte copy: 0 ULP for any element type.  Asynchronous on `stream`.
 * `variant`: 0 = auto, 1 = batched SIMT row gather, 2 = TMA bulk-copy pipeline (needs row_bytes % 16 == 0),
 *            3 = flat chunked SIMT gather (kept for comparison). */
QV_API int qv_gather(const qv_shard_table *table, const int64_t *indices, const int64_t *feature_order, int64_t n,
              int64_t row_bytes, void *out, int variant, qv_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * CSR neighbour sampler
 * ------------------------------------------------------------------------------------------------------- */

typedef struct qv_sampler qv_sampler;

/* torch_quiver.device_quiver_from_csr_array(indptr, indices, edge_ids, device, cuda)
 * -- srcs/cpp/src/quiver/cuda/quiver_sample.cu:361-461.  `indptr` ([n_nodes+1]) and `indices` ([n_edges]) are
 * pointers VALID ON `device`: HBM copies for mode="GPU", a qv_host_register alias of the caller's CPU tensor for
 * mode="UVA" (the caller keeps that tensor alive, as in the reference).  The object borrows both arrays and owns
 * its scratch (hash table, scan state, XORWOW state cache). */
QV_API int qv_sampler_create(int device, const int64_t *indptr, int64_t n_nodes, const int64_t *indices, int64_t n_edges,
                      qv_sampler **out);
QV_API int qv_sampler_destroy(qv_sampler *s);

/* The `edge_ids` argument of device_quiver_from_csr_array (quiver_sample.cu:434-453: one id per CSR position, HBM copy
 * or registered host alias; the reference stores it and every sampler then returns an empty e_id,
 * sage_sampler.py:143).  Here it is used: the edge-id outputs below emit edge_ids[p] for a sampled CSR position p, or
 * p itself while no array is set (NULL).  Borrowed, [n_edges] int64 valid on the sampler's device. */
QV_API int qv_sampler_set_edge_ids(qv_sampler *s, const int64_t *edge_ids);

/* Quiver.sample_neighbor, first half -- quiver_sample.cu:157-169 (degree, cap_by(k), exclusive_scan, reduce).
 *   counts[i] = min(deg(seeds[i]), k)   (k < 0: no cap; a seed outside [0, n_nodes) counts as degree 0)
 *   out_ptr   = exclusive_scan(counts);  *total = sum(counts)
 * counts / out_ptr are device arrays of S int64.  Synchronises `stream` once to return *total. */
QV_API int qv_sample_count(qv_sampler *s, const int64_t *seeds, int64_t S, int64_t k, int64_t *counts, int64_t *out_ptr,
                    int64_t *total, qv_stream_t stream);

/* Quiver.sample_neighbor, second half -- CSRRowWiseSampleKernel (include/quiver/cuda_random.cu.hpp:7-69) as
 * launched by quiver<T,CUDA>::new_sample (include/quiver/quiver.cu.hpp:380-403).  Writes, for every seed i, its
 * min(deg,k) sampled neighbour ids to neighbors[out_ptr[i] ...]: the row verbatim when deg <= k, otherwise
 * exactly the ids the reference kernel picks for generator seed `rand_seed` (the reference hard-codes 0).
 * edge_ids_out (optional, same layout as neighbors): the edge id of every sampled neighbour -- its CSR position
 * indptr[seed] + pos, mapped through qv_sampler_set_edge_ids when set (the reference's sample kernel has
 * the plumbing, quiver.cu.hpp:90-126, and drops the result).  Asynchronous. */
QV_API int qv_sample_fill(qv_sampler *s, const int64_t *seeds, int64_t S, int64_t k, uint64_t rand_seed,
                   const int64_t *out_ptr, int64_t *neighbors, int64_t *edge_ids_out, qv_stream_t stream);

/* Quiver.reindex_single(inputs, outputs, counts) -- quiver_sample.cu:305-357 (reindex_kernel :202-255,
 * FillWithDuplicates :18-63, DeviceOrderedHashTable include/quiver/reindex.cu.hpp:20-158).
 *   frontier = unique(concat(inputs, outputs)) in first-occurrence order   (capacity S + tot int64)
 *   col_idx[e] = position of outputs[e] in frontier;  row_idx[e] = index of the seed that produced e
 * Synchronises `stream` once to return *n_frontier. */
QV_API int qv_reindex(qv_sampler *s, const int64_t *inputs, int64_t S, const int64_t *outputs, int64_t tot,
               const int64_t *counts, int64_t *frontier, int64_t *row_idx, int64_t *col_idx, int64_t *n_frontier,
               qv_stream_t stream);

/* GraphSageSampler.sample(seeds) -- srcs/python/quiver/pyg/sage_sampler.py:118-147, all hops in one call with a
 * single stream synchronisation at the end (the reference synchronises ~3 times per hop).
 * Hop l (0-based) samples sizes[l] neighbours of frontier l (frontier 0 = seeds) and reindexes.
 * Buffers are caller-allocated to the upper bounds returned by qv_khop_bounds:
 *   n_id      [bound_nodes[n_hops]]       final frontier (n_id of the PyG triple), valid prefix out_nodes[n_hops]
 *   edge_buf[l] [2 * bound_edges[l]]      hop l's edge_index, stored as two back-to-back rows of out_edges[l]:
 *                                         [0,E) = source (neighbour) local ids, [E,2E) = target (seed) local ids,
 *                                         i.e. buf[:2E].view(2, E) is the contiguous PyG edge_index
 *   eid_buf   NULL, or [n_hops] pointers (each NULL or [bound_edges[l]]): edge id of hop l's e-th edge (PyG's e_id;
 *                                         see qv_sample_fill), valid prefix out_edges[l]
 *   out_nodes [n_hops+1], out_edges [n_hops]   host arrays: |frontier l| (out_nodes[0] = S) and E_l
 * All sizes[l] must be >= 0 (use the per-hop calls for "-1 = all neighbours").
 * With n_hops = 1 this is Quiver.sample_sub(stream_num, vertices, k) -- quiver_sample.cu:257-304 -- as ONE call:
 * frontier = n_id, col_idx = edge_buf[0][0:E], row_idx = edge_buf[0][E:2E]. */
QV_API int qv_khop_bounds(int64_t S, const int64_t *sizes, int n_hops, int64_t *bound_nodes /* [n_hops+1] */,
                   int64_t *bound_edges /* [n_hops] */);
QV_API int qv_khop(qv_sampler *s, const int64_t *seeds, int64_t S, const int64_t *sizes, int n_hops, uint64_t rand_seed,
            int64_t *n_id, int64_t *const *edge_buf, int64_t *const *eid_buf, int64_t *out_nodes, int64_t *out_edges,
            qv_stream_t stream);

/* sample -> gather without returning to the host in between (the two reference calls it replaces are
 * GraphSageSampler.sample, sage_sampler.py:118-147, followed by Feature.__getitem__(n_id), feature.py:296-308).
 * Same outputs as qv_khop, plus features[i, :] = row feature_order[n_id[i]] of `table` for i < out_nodes[n_hops]
 * (qv_gather semantics; feature_order may be NULL).  The gather is enqueued BEHIND the sampling kernels with the
 * frontier size read on the device, and the host wait covers only the sampler's sizes: the call returns while the
 * gather may still be running on `stream` (stream-ordered like qv_gather).
 *   features  [features_rows * row_bytes]   caller-allocated; valid prefix out_nodes[n_hops] rows.  features_rows =
 *             bound_nodes[n_hops] always suffices; a frontier holds distinct nodes, so min(bound, n_nodes + S) does too
 *             (the adapter uses that: on graphs smaller than the fan-out bound it is many times smaller).  If the
 *             frontier turns out larger than features_rows the first features_rows rows are gathered, the sample
 *             outputs are complete and the call returns QV_ERR_UNSUPPORTED.
 * The table must be usable from the sampler's device. */
QV_API int qv_khop_gather(qv_sampler *s, const int64_t *seeds, int64_t S, const int64_t *sizes, int n_hops,
                   uint64_t rand_seed, int64_t *n_id, int64_t *const *edge_buf, int64_t *const *eid_buf,
                   const struct qv_shard_table *table, const int64_t *feature_order, int64_t row_bytes, void *features,
                   int64_t features_rows, int variant, int64_t *out_nodes, int64_t *out_edges, qv_stream_t stream);

/* Quiver.cal_neighbor_prob(stream_num, last_prob, cur_prob, k) -- quiver_sample.cu:100-111 launching cal_next
 * (include/quiver/cuda_random.cu.hpp:71-104): one hop of access-probability propagation, fp32, same operation
 * order per node.  Asynchronous. */
QV_API int qv_cal_neighbor_prob(qv_sampler *s, const float *last_prob, float *cur_prob, int64_t n, int k,
                         qv_stream_t stream);

/* Extension (no reference counterpart): opt into position-independent O(k)-per-row sampling.  Same contract as
 * CSRRowWiseSampleKernel (min(deg,k) distinct uniform positions, verbatim copy when deg <= k) but NOT the reference's
 * random stream: ids differ from the reference's, and each call draws a fresh sample (seed = rand_seed + call counter).
 * Default 0 = bit-identical to the reference under its generator seed. */
QV_API int qv_sampler_set_fast(qv_sampler *s, int enabled);

/* Diagnostics: number of kernel launches this library has issued in this process (bench.py "gpu_launches"). */
QV_API int64_t qv_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QUIVER_B200_H */
