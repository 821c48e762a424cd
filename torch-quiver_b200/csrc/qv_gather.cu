// qv_gather.cu -- tiered / sharded feature gather ("collect") for B200 (sm_100a).
//
// Replaces quiver_tensor_gather + find (srcs/cpp/include/quiver/shard_tensor.cu.hpp:7-61) as launched by
// ShardTensor::operator[] (srcs/cpp/src/quiver/cuda/quiver_feature.cu:246-302), and folds in the
// feature_order[idx] indirection of Feature.__getitem__ (srcs/python/quiver/feature.py:300-301).
//
// The reference moves ONE BYTE per lane per instruction (32 B per warp instruction) and re-reads the shard offset
// table from global memory for every row.  Here:
//   * the shard table (<= 16 entries) travels in the kernel parameters (__grid_constant__, constant bank reads);
//   * variant 1 (SIMT): the output is treated as a flat array of 16-byte chunks; a thread moves `kUnroll` chunks with
//     all loads issued before the first store (MLP), 128-bit streaming loads from whichever tier owns the row
//     (local HBM / peer HBM over NVLink / pinned host over PCIe -- the pointer decides) and perfectly coalesced
//     128-bit streaming stores.  Rows whose size is not a multiple of 16 fall back to 8/4/2/1-byte chunks;
//   * variant 2 (TMA): rows are pulled with cp.async.bulk (one bulk copy per row, issued by single lanes, completion
//     on an mbarrier) into a shared-memory ring and pushed out with ONE bulk store per stage, because consecutive
//     output rows are contiguous.  No registers touch the payload.  Needs row_bytes % 16 == 0.
//   * ids outside [0, rows) and inaccessible shards produce zero rows (the reference leaves them uninitialised).
#include <algorithm>
#include <cstdlib>

#include "qv_common.cuh"

namespace qv
{
namespace
{
struct GatherParams {
    int32_t n_shards;
    int32_t _pad;
    int64_t row_begin[QV_MAX_SHARDS + 1];
    const char *ptr[QV_MAX_SHARDS];  // nullptr = not accessible from this device
    int64_t pitch[QV_MAX_SHARDS];
};

// Source address of logical row `id`, or nullptr for "zero row".
__device__ __forceinline__ const char *row_source(const GatherParams &t, int64_t id)
{
    if (id < 0 || id >= t.row_begin[t.n_shards]) return nullptr;
    int s = 0;
#pragma unroll 4
    for (int i = 1; i < QV_MAX_SHARDS; i++)
        if (i < t.n_shards && id >= t.row_begin[i]) s = i;
    const char *base = t.ptr[s];
    return base ? base + (id - t.row_begin[s]) * t.pitch[s] : nullptr;
}

__device__ __forceinline__ int64_t logical_row(const int64_t *__restrict__ indices,
                                               const int64_t *__restrict__ feature_order, int64_t n_rows_total,
                                               int64_t i)
{
    int64_t id = indices[i];
    if (feature_order) id = (id >= 0 && id < n_rows_total) ? feature_order[id] : -1;
    return id;
}

// Number of rows to gather: `n` from the host, capped by a device-resident count when the caller does not know the
// size on the host yet (qv_khop_gather: the frontier size is still being computed when the gather is enqueued; `n` is
// then the static bound the grid was sized for).
__device__ __forceinline__ int64_t live_rows(int64_t n, const int64_t *__restrict__ d_n)
{
    if (d_n) n = min(n, max(static_cast<int64_t>(0), *d_n));
    return n;
}

template <int kBytes>
struct Chunk;
template <>
struct Chunk<16> {
    using T = int4;
    static __device__ __forceinline__ T load(const char *p) { return ld_stream_v4(p); }
    static __device__ __forceinline__ void store(char *p, const T &v) { st_stream_v4(p, v); }
    static __device__ __forceinline__ T zero() { return make_int4(0, 0, 0, 0); }
};
template <>
struct Chunk<8> {
    using T = int2;
    static __device__ __forceinline__ T load(const char *p) { return ld_stream_v2(p); }
    static __device__ __forceinline__ void store(char *p, const T &v) { st_stream_v2(p, v); }
    static __device__ __forceinline__ T zero() { return make_int2(0, 0); }
};
template <>
struct Chunk<4> {
    using T = int;
    static __device__ __forceinline__ T load(const char *p) { return __ldg(reinterpret_cast<const int *>(p)); }
    static __device__ __forceinline__ void store(char *p, const T &v) { *reinterpret_cast<int *>(p) = v; }
    static __device__ __forceinline__ T zero() { return 0; }
};
template <>
struct Chunk<2> {
    using T = short;
    static __device__ __forceinline__ T load(const char *p) { return __ldg(reinterpret_cast<const short *>(p)); }
    static __device__ __forceinline__ void store(char *p, const T &v) { *reinterpret_cast<short *>(p) = v; }
    static __device__ __forceinline__ T zero() { return 0; }
};
template <>
struct Chunk<1> {
    using T = char;
    static __device__ __forceinline__ T load(const char *p) { return __ldg(p); }
    static __device__ __forceinline__ void store(char *p, const T &v) { *p = v; }
    static __device__ __forceinline__ T zero() { return 0; }
};

// ------------------------------------------------------------------------------------------------------------------
// Variant 1: flat chunked SIMT gather.
// Block = kThreads threads, each block owns kThreads*kUnroll consecutive output chunks.  chunks_per_row (cpr) and the
// multiplier inv = floor(2^32/cpr)+1 give an exact 32-bit division for the block-local chunk offset (< cpr + tile).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kGatherUnroll = 4;

template <int kBytes>
__global__ void __launch_bounds__(kGatherThreads)
    gather_flat_kernel(const __grid_constant__ GatherParams t, const int64_t *__restrict__ indices,
                       const int64_t *__restrict__ feature_order, int64_t n, const int64_t *__restrict__ d_n,
                       uint32_t cpr, uint32_t inv, char *__restrict__ out)
{
    using C = Chunk<kBytes>;
    constexpr uint32_t kTile = kGatherThreads * kGatherUnroll;
    n = live_rows(n, d_n);
    const int64_t total_chunks = n * static_cast<int64_t>(cpr);
    const int64_t tile_base = static_cast<int64_t>(blockIdx.x) * kTile;
    // first row of the tile and the tile's chunk offset inside that row (one 64-bit division per thread)
    const int64_t row0 = tile_base / cpr;
    const uint32_t off0 = static_cast<uint32_t>(tile_base - row0 * cpr);
    const int64_t n_rows_total = t.row_begin[t.n_shards];

    typename C::T v[kGatherUnroll];
    const char *src[kGatherUnroll];
    bool live[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; u++) {
        const uint32_t local = threadIdx.x + u * kGatherThreads;
        live[u] = tile_base + local < total_chunks;
        src[u] = nullptr;
        if (live[u]) {
            const uint32_t l = off0 + local;
            const uint32_t dr = cpr == 1 ? l : __umulhi(l, inv);  // exact: l * cpr < 2^32 (checked by the host)
            const uint32_t col = l - dr * cpr;
            const int64_t id = logical_row(indices, feature_order, n_rows_total, row0 + dr);
            const char *base = row_source(t, id);
            if (base) src[u] = base + static_cast<size_t>(col) * kBytes;
        }
    }
#pragma unroll
    for (int u = 0; u < kGatherUnroll; u++) v[u] = src[u] ? C::load(src[u]) : C::zero();
#pragma unroll
    for (int u = 0; u < kGatherUnroll; u++) {
        if (live[u]) {
            const int64_t c = tile_base + threadIdx.x + u * kGatherThreads;
            C::store(out + c * kBytes, v[u]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Variant 1 (default): batched row gather.  A warp owns 32 consecutive output rows.
//   phase 1  lane r resolves row r: coalesced 8-byte index load, optional feature_order hop, shard lookup -> source
//            pointer (32 independent dependent-load chains in flight per warp);
//   phase 2  the rows are copied kUnroll at a time: the source pointer of a row is broadcast with shuffles, lanes move
//            16-byte chunks (all loads of the kUnroll rows are issued before the first store).  Rows shorter than a
//            warp-width of chunks are packed kGroup lanes per row so no lane idles.
// kLoad/kStore: 16/16 when everything is 16-byte aligned; 16/8 when rows are 8-byte multiples but sources are padded
// to 16 (e.g. 602 fp32 = 2408 B rows: 16-byte loads, split stores); otherwise kLoad == kStore in {8,4,2,1}.
// About 10 instructions per row-chunk-iteration instead of ~190 for the flat variant (no division, one shard search
// per row instead of per chunk).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBatchUnroll = 4;

template <int kLoad, int kStore, int kGroup>
__global__ void __launch_bounds__(256)
    gather_batch_kernel(const __grid_constant__ GatherParams t, const int64_t *__restrict__ indices,
                        const int64_t *__restrict__ feature_order, int64_t n, const int64_t *__restrict__ d_n,
                        uint32_t row_bytes, char *__restrict__ out)
{
    using L = Chunk<kLoad>;
    constexpr int kRowsPerIter = 32 / kGroup;
    n = live_rows(n, d_n);
    const int lane = threadIdx.x & 31;
    const int sub = lane % kGroup, grp = lane / kGroup;
    const int64_t n_rows_total = t.row_begin[t.n_shards];
    const uint32_t cpr = (row_bytes + kLoad - 1) / kLoad;  // load-chunks per row (last one may be half used: 16/8 mode)
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t base = warp * 32;
    if (base >= n) return;

    const char *my_src = nullptr;
    if (base + lane < n) my_src = row_source(t, logical_row(indices, feature_order, n_rows_total, base + lane));
    const int rows_here = static_cast<int>(min(static_cast<int64_t>(32), n - base));
    char *out_base = out + base * row_bytes;

    for (int rr = 0; rr < rows_here; rr += kRowsPerIter * kBatchUnroll) {
        const char *src[kBatchUnroll];
        int row[kBatchUnroll];
#pragma unroll
        for (int u = 0; u < kBatchUnroll; u++) {
            row[u] = rr + u * kRowsPerIter + grp;
            const unsigned long long p = __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(my_src), row[u] & 31);
            src[u] = reinterpret_cast<const char *>(p);
        }
        for (uint32_t c = sub; c < cpr; c += kGroup) {
            typename L::T v[kBatchUnroll];
#pragma unroll
            for (int u = 0; u < kBatchUnroll; u++)
                v[u] = (row[u] < rows_here && src[u]) ? L::load(src[u] + static_cast<size_t>(c) * kLoad) : L::zero();
#pragma unroll
            for (int u = 0; u < kBatchUnroll; u++) {
                if (row[u] < rows_here) {
                    char *dst = out_base + static_cast<size_t>(row[u]) * row_bytes + static_cast<size_t>(c) * kLoad;
                    if constexpr (kLoad == kStore) {
                        L::store(dst, v[u]);
                    } else {  // 16-byte load, two 8-byte stores; the row's last chunk may hold only 8 valid bytes
                        static_assert(kLoad == 16 && kStore == 8, "only the 16/8 split is implemented");
                        st_stream_v2(dst, make_int2(v[u].x, v[u].y));
                        if (c * 16u + 8u < row_bytes) st_stream_v2(dst + 8, make_int2(v[u].z, v[u].w));
                    }
                }
            }
        }
    }
}

// Same batching, but the 32 rows of a warp are walked as ONE flat run of chunks (chunk f -> row f / cpr, column
// f % cpr, exact 32-bit multiply-high division), so rows whose chunk count is not a multiple of the group width keep
// every lane busy: 400-byte rows (25 chunks) use 32/32 lanes instead of 25/32.
template <int kLoad, int kStore, int kUnroll = kBatchUnroll, int kMinBlocks = 5>
__global__ void __launch_bounds__(256, kMinBlocks)
    gather_batch_flat_kernel(const __grid_constant__ GatherParams t, const int64_t *__restrict__ indices,
                             const int64_t *__restrict__ feature_order, int64_t n, const int64_t *__restrict__ d_n,
                             uint32_t row_bytes, uint32_t cpr, uint32_t inv, char *__restrict__ out)
{
    using L = Chunk<kLoad>;
    const int lane = threadIdx.x & 31;
    n = live_rows(n, d_n);
    const int64_t n_rows_total = t.row_begin[t.n_shards];
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t base = warp * 32;
    if (base >= n) return;

    const unsigned long long pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const char *my_src = nullptr;
    if (base + lane < n) {
        long long id = indices[base + lane];
        if (feature_order) id = (id >= 0 && id < n_rows_total) ? ld_keep_s64(feature_order + id, pol_keep) : -1;
        my_src = row_source(t, id);
    }
    const uint32_t rows_here = static_cast<uint32_t>(min(static_cast<int64_t>(32), n - base));
    const uint32_t total = rows_here * cpr;
    char *out_base = out + base * row_bytes;

    for (uint32_t f0 = 0; f0 < total; f0 += 32 * kUnroll) {
        const char *src[kUnroll];
        uint32_t off[kUnroll];  // byte offset of the chunk inside the warp's output run
        uint32_t col[kUnroll];
        bool live[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const uint32_t f = f0 + u * 32 + lane;
            live[u] = f < total;
            const uint32_t row = live[u] ? __umulhi(f, inv) : 0;  // exact: f * cpr < 2^32 (host-checked)
            col[u] = f - row * cpr;
            off[u] = row * row_bytes + col[u] * kLoad;
            const unsigned long long p = __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(my_src), row);
            src[u] = reinterpret_cast<const char *>(p);
        }
        typename L::T v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
        {
            if constexpr (kLoad == 16)
                v[u] = (live[u] && src[u]) ? ld_stream_v4_hint(src[u] + static_cast<size_t>(col[u]) * kLoad, pol_stream)
                                           : L::zero();
            else
                v[u] = (live[u] && src[u]) ? L::load(src[u] + static_cast<size_t>(col[u]) * kLoad) : L::zero();
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            if (live[u]) {
                char *dst = out_base + off[u];
                if constexpr (kLoad == 16 && kStore == 16) {
                    st_stream_v4_hint(dst, v[u], pol_stream);
                } else if constexpr (kLoad == kStore) {
                    L::store(dst, v[u]);
                } else {
                    static_assert(kLoad == 16 && kStore == 8, "only the 16/8 split is implemented");
                    st_stream_v2(dst, make_int2(v[u].x, v[u].y));
                    if (col[u] * 16u + 8u < row_bytes) st_stream_v2(dst + 8, make_int2(v[u].z, v[u].w));
                }
            }
        }
    }
}

// Fallback for rows too long for the 32-bit trick: one warp per row, 64-bit addressing.
template <int kBytes>
__global__ void __launch_bounds__(256)
    gather_rows_kernel(const __grid_constant__ GatherParams t, const int64_t *__restrict__ indices,
                       const int64_t *__restrict__ feature_order, int64_t n, const int64_t *__restrict__ d_n,
                       int64_t cpr, char *__restrict__ out)
{
    using C = Chunk<kBytes>;
    n = live_rows(n, d_n);
    const int64_t n_rows_total = t.row_begin[t.n_shards];
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += n_warps) {
        const char *base = row_source(t, logical_row(indices, feature_order, n_rows_total, r));
        char *dst = out + r * cpr * kBytes;
        for (int64_t c = lane; c < cpr; c += 32) C::store(dst + c * kBytes, base ? C::load(base + c * kBytes) : C::zero());
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Variant 2: TMA bulk-copy pipeline.  One CTA = one producer warp; a stage holds kRowsPerStage consecutive output
// rows.  Lane l issues the bulk load of row l of the stage (global -> shared, completes on the stage's mbarrier);
// when the barrier flips, lane 0 issues one bulk store of the whole stage (shared -> global) and the ring advances.
// Zero rows are written into shared memory by the lanes themselves.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTmaStages = 6;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_addr(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kKeep>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kKeep) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(32)
    gather_tma_kernel(const __grid_constant__ GatherParams t, const int64_t *__restrict__ indices,
                      const int64_t *__restrict__ feature_order, int64_t n, const int64_t *__restrict__ d_n,
                      uint32_t row_bytes, int rows_per_stage, char *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    n = live_rows(n, d_n);
    __shared__ __align__(8) uint64_t full[kTmaStages];
    const int lane = threadIdx.x;
    const uint32_t stage_bytes = row_bytes * rows_per_stage;
    const int64_t n_rows_total = t.row_begin[t.n_shards];
    if (lane == 0) {
        for (int s = 0; s < kTmaStages; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int64_t n_groups = (n + rows_per_stage - 1) / rows_per_stage;
    // this CTA handles groups blockIdx.x, blockIdx.x + gridDim.x, ...
    int64_t issue = blockIdx.x;  // next group to load
    int64_t drain = blockIdx.x;  // next group to store
    int issue_slot = 0, drain_slot = 0;
    uint32_t phase_bits = 0;  // bit s = parity to wait for on stage s

    auto issue_group = [&](int64_t g, int slot) {
        unsigned char *stage = smem_raw + static_cast<size_t>(slot) * stage_bytes;
        const int64_t first = g * rows_per_stage;
        const int rows = static_cast<int>(min(static_cast<int64_t>(rows_per_stage), n - first));
        // expected bytes = rows that really come from memory; zero rows are filled by hand
        uint32_t tx = 0;
        const char *src[4];
        for (int it = 0; it * 32 < rows_per_stage; it++) {
            const int r = lane + it * 32;
            src[it] = nullptr;
            if (r < rows) src[it] = row_source(t, logical_row(indices, feature_order, n_rows_total, first + r));
            tx += __popc(__ballot_sync(0xffffffffu, src[it] != nullptr)) * row_bytes;
            if (r < rows && src[it] == nullptr) {
                int4 *z = reinterpret_cast<int4 *>(stage + static_cast<size_t>(r) * row_bytes);
                for (uint32_t c = 0; c < row_bytes / 16; c++) z[c] = make_int4(0, 0, 0, 0);
            }
        }
        fence_proxy_async();  // hand-written zero rows must be visible to the bulk store
        __syncwarp();
        if (lane == 0) mbar_expect_tx(&full[slot], tx);
        __syncwarp();
        for (int it = 0; it * 32 < rows_per_stage; it++) {
            const int r = lane + it * 32;
            if (src[it]) bulk_load(stage + static_cast<size_t>(r) * row_bytes, src[it], row_bytes, &full[slot]);
        }
    };

    // prologue: fill the ring
    for (int s = 0; s < kTmaStages && issue < n_groups; s++) {
        issue_group(issue, issue_slot);
        issue += gridDim.x;
        issue_slot = (issue_slot + 1) % kTmaStages;
    }
    int prev_slot = -1;
    while (drain < n_groups) {
        mbar_wait(&full[drain_slot], (phase_bits >> drain_slot) & 1u);
        phase_bits ^= 1u << drain_slot;
        const int64_t first = drain * rows_per_stage;
        const int rows = static_cast<int>(min(static_cast<int64_t>(rows_per_stage), n - first));
        if (lane == 0) {
            bulk_store(out + first * row_bytes, smem_raw + static_cast<size_t>(drain_slot) * stage_bytes,
                       static_cast<uint32_t>(rows) * row_bytes);
            bulk_commit();
        }
        drain += gridDim.x;
        // refill the slot drained one trip ago: its store has finished READING shared memory once at most one store
        // group (the one just committed) is still pending
        if (prev_slot >= 0 && issue < n_groups) {
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            issue_group(issue, prev_slot);
            issue += gridDim.x;
        }
        prev_slot = drain_slot;
        drain_slot = (drain_slot + 1) % kTmaStages;
    }
    if (lane == 0) bulk_wait_read<0>();
    (void)issue_slot;
}

inline int pick_chunk(int64_t row_bytes, const qv_shard_table *tab, const void *out)
{
    uintptr_t bits = static_cast<uintptr_t>(row_bytes) | reinterpret_cast<uintptr_t>(out);
    for (int s = 0; s < tab->n_shards; s++) {
        bits |= reinterpret_cast<uintptr_t>(tab->ptr[s]);
        bits |= static_cast<uintptr_t>(tab->pitch[s]);
    }
    if ((bits & 15) == 0) return 16;
    if ((bits & 7) == 0) return 8;
    if ((bits & 3) == 0) return 4;
    if ((bits & 1) == 0) return 2;
    return 1;
}

template <int kLoad, int kStore>
int launch_batch(const GatherParams &p, const int64_t *indices, const int64_t *feature_order, int64_t n,
                 const int64_t *d_n, int64_t row_bytes, char *out, cudaStream_t st)
{
    const int64_t cpr = (row_bytes + kLoad - 1) / kLoad;
    const int64_t warps = (n + 31) / 32;
    const int64_t blocks = (warps + 7) / 8;
    QV_REQUIRE(blocks < (int64_t(1) << 31) && row_bytes < (int64_t(1) << 31), "qv_gather: request too large");
    const unsigned g = static_cast<unsigned>(blocks);
    const uint32_t rb = static_cast<uint32_t>(row_bytes);
    if ((cpr & (cpr - 1)) != 0 && (cpr % 32) != 0 && cpr <= 8192 && row_bytes * 32 < (int64_t(1) << 31)) {
        // rows that would leave lanes idle in the grouped kernel: flat walk (32 * cpr * cpr < 2^32 holds)
        const uint32_t inv = static_cast<uint32_t>((uint64_t(1) << 32) / static_cast<uint64_t>(cpr)) + 1u;
        // (rows in flight per lane, min blocks per SM) = (4, 5): a sweep over {2,4,8} x {3..8} stayed within 0.75-0.79
        // of the HBM peak on 400-byte rows -- the kernel sits on the DRAM random-access limit, not on occupancy
        gather_batch_flat_kernel<kLoad, kStore><<<g, 256, 0, st>>>(p, indices, feature_order, n, d_n, rb,
                                                                    static_cast<uint32_t>(cpr), inv, out);
        QV_CHECK_LAUNCH("gather_batch_flat_kernel");
        return QV_OK;
    }
#define QV_LAUNCH_GROUP(G)                                                                                     \
    gather_batch_kernel<kLoad, kStore, G><<<g, 256, 0, st>>>(p, indices, feature_order, n, d_n, rb, out)
    if (cpr > 16)
        QV_LAUNCH_GROUP(32);
    else if (cpr > 8)
        QV_LAUNCH_GROUP(16);
    else if (cpr > 4)
        QV_LAUNCH_GROUP(8);
    else if (cpr > 2)
        QV_LAUNCH_GROUP(4);
    else if (cpr > 1)
        QV_LAUNCH_GROUP(2);
    else
        QV_LAUNCH_GROUP(1);
#undef QV_LAUNCH_GROUP
    QV_CHECK_LAUNCH("gather_batch_kernel");
    return QV_OK;
}

template <int kBytes>
int launch_simt(const GatherParams &p, const int64_t *indices, const int64_t *feature_order, int64_t n,
                const int64_t *d_n, int64_t row_bytes, char *out, int n_sm, cudaStream_t st)
{
    const int64_t cpr = row_bytes / kBytes;
    constexpr int64_t kTile = kGatherThreads * kGatherUnroll;
    const int64_t total = n * cpr;
    const int64_t blocks = (total + kTile - 1) / kTile;
    // exactness of the __umulhi division needs (cpr + tile) * cpr < 2^32
    if ((cpr + kTile) * cpr < (int64_t(1) << 32) && blocks < (int64_t(1) << 31)) {
        const uint32_t inv = static_cast<uint32_t>((uint64_t(1) << 32) / static_cast<uint64_t>(cpr)) + 1u;
        gather_flat_kernel<kBytes><<<static_cast<unsigned>(blocks), kGatherThreads, 0, st>>>(
            p, indices, feature_order, n, d_n, static_cast<uint32_t>(cpr), cpr == 1 ? 0u : inv, out);
        QV_CHECK_LAUNCH("gather_flat_kernel");
    } else {
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n + 7) / 8, int64_t(n_sm) * 16));
        gather_rows_kernel<kBytes><<<grid, 256, 0, st>>>(p, indices, feature_order, n, d_n, cpr, out);
        QV_CHECK_LAUNCH("gather_rows_kernel");
    }
    return QV_OK;
}
}  // namespace

// Validates the request and enqueues the gather.  d_n (optional) = device-resident row count, n = its host-side bound.
int gather_enqueue(const qv_shard_table *table, const int64_t *indices, const int64_t *feature_order, int64_t n,
                   const int64_t *d_n, int64_t row_bytes, void *out, int variant, cudaStream_t st)
{
    QV_REQUIRE(table != nullptr, "qv_gather: table is NULL");
    QV_REQUIRE(table->n_shards >= 1 && table->n_shards <= QV_MAX_SHARDS, "qv_gather: n_shards = %d outside [1, %d]",
               table->n_shards, QV_MAX_SHARDS);
    QV_REQUIRE(n >= 0 && row_bytes > 0, "qv_gather: bad sizes (n = %lld, row_bytes = %lld)", (long long)n,
               (long long)row_bytes);
    if (n == 0) return QV_OK;
    QV_REQUIRE(indices && out, "qv_gather: NULL array");
    GatherParams p;
    memset(&p, 0, sizeof p);
    p.n_shards = table->n_shards;
    QV_REQUIRE(table->row_begin[0] == 0, "qv_gather: row_begin[0] must be 0");
    for (int s = 0; s < table->n_shards; s++) {
        QV_REQUIRE(table->row_begin[s + 1] >= table->row_begin[s], "qv_gather: row_begin not monotone at shard %d", s);
        QV_REQUIRE(table->pitch[s] >= row_bytes || table->row_begin[s + 1] == table->row_begin[s],
                   "qv_gather: shard %d pitch %lld < row_bytes %lld", s, (long long)table->pitch[s],
                   (long long)row_bytes);
        p.row_begin[s] = table->row_begin[s];
        p.ptr[s] = (table->accessible[s] && table->ptr[s]) ? static_cast<const char *>(table->ptr[s]) : nullptr;
        p.pitch[s] = table->pitch[s];
    }
    for (int s = table->n_shards; s <= QV_MAX_SHARDS; s++) p.row_begin[s] = table->row_begin[table->n_shards];
    p.row_begin[table->n_shards] = table->row_begin[table->n_shards];

    int device = 0;
    QV_CUDA(cudaGetDevice(&device));
    const int n_sm = sm_count(device);
    char *o = static_cast<char *>(out);
    const int chunk = pick_chunk(row_bytes, table, out);

    // auto: rows of >= 2 KiB go through the TMA pipeline (measured equal to the SIMT kernel at 3 KiB rows, 0.93 of the HBM
    // peak, with no payload in registers); shorter rows are faster with 16-byte SIMT accesses (0.80 vs 0.67 at 400 B)
    if (variant == 2 || (variant == 0 && chunk == 16 && row_bytes >= 2048 && row_bytes <= 8 * 1024)) {
        QV_REQUIRE(chunk == 16, "qv_gather: the TMA variant needs 16-byte aligned rows, pitches and pointers");
        QV_REQUIRE(row_bytes <= 32 * 1024, "qv_gather: the TMA variant supports rows up to 32 KiB");
        // stage = up to 32 rows of ~8 KiB in total; 6 stages per CTA, several CTAs per SM
        int rows_per_stage = static_cast<int>(std::min<int64_t>(32, (8 * 1024) / row_bytes));
        rows_per_stage = std::max(rows_per_stage, 1);
        const size_t smem = static_cast<size_t>(kTmaStages) * rows_per_stage * row_bytes;
        QV_REQUIRE(smem <= 200 * 1024, "qv_gather: the TMA variant stages %d rows x %d buffers: rows above %d bytes do not fit",
                   rows_per_stage, kTmaStages, 200 * 1024 / kTmaStages);
        static std::atomic<unsigned long long> attr_set{0};  // bit d: the attribute is set on device d (any thread may race here: setting it twice is harmless)
        const unsigned long long bit = (device >= 0 && device < 64) ? 1ull << device : 0ull;
        if (!(attr_set.load(std::memory_order_acquire) & bit) || bit == 0) {
            QV_CUDA(cudaFuncSetAttribute(gather_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set.fetch_or(bit, std::memory_order_release);
        }
        const int64_t n_groups = (n + rows_per_stage - 1) / rows_per_stage;
        const int ctas_per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(16, (200 * 1024) / (smem + 1024))));
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>(n_groups, int64_t(n_sm) * ctas_per_sm));
        gather_tma_kernel<<<grid, 32, smem, st>>>(p, indices, feature_order, n, d_n, static_cast<uint32_t>(row_bytes),
                                                   rows_per_stage, o);
        QV_CHECK_LAUNCH("gather_tma_kernel");
        return QV_OK;
    }
    if (variant != 3) {
        // can sources be read 16 bytes at a time even though rows are only 8-byte multiples?
        bool src16 = true;
        for (int s = 0; s < table->n_shards; s++)
            if (table->row_begin[s + 1] > table->row_begin[s] &&
                ((reinterpret_cast<uintptr_t>(table->ptr[s]) | static_cast<uintptr_t>(table->pitch[s])) & 15))
                src16 = false;
        switch (chunk) {
        case 16:
            return launch_batch<16, 16>(p, indices, feature_order, n, d_n, row_bytes, o, st);
        case 8:
            if (src16) return launch_batch<16, 8>(p, indices, feature_order, n, d_n, row_bytes, o, st);
            return launch_batch<8, 8>(p, indices, feature_order, n, d_n, row_bytes, o, st);
        case 4:
            return launch_batch<4, 4>(p, indices, feature_order, n, d_n, row_bytes, o, st);
        case 2:
            return launch_batch<2, 2>(p, indices, feature_order, n, d_n, row_bytes, o, st);
        default:
            return launch_batch<1, 1>(p, indices, feature_order, n, d_n, row_bytes, o, st);
        }
    }
    switch (chunk) {
    case 16:
        return launch_simt<16>(p, indices, feature_order, n, d_n, row_bytes, o, n_sm, st);
    case 8:
        return launch_simt<8>(p, indices, feature_order, n, d_n, row_bytes, o, n_sm, st);
    case 4:
        return launch_simt<4>(p, indices, feature_order, n, d_n, row_bytes, o, n_sm, st);
    case 2:
        return launch_simt<2>(p, indices, feature_order, n, d_n, row_bytes, o, n_sm, st);
    default:
        return launch_simt<1>(p, indices, feature_order, n, d_n, row_bytes, o, n_sm, st);
    }
}
}  // namespace qv

using namespace qv;

extern "C" int qv_gather(const qv_shard_table *table, const int64_t *indices, const int64_t *feature_order, int64_t n,
                         int64_t row_bytes, void *out, int variant, qv_stream_t stream)
{
    return gather_enqueue(table, indices, feature_order, n, nullptr, row_bytes, out, variant,
                          static_cast<cudaStream_t>(stream));
}
