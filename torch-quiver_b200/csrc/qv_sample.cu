// qv_sample.cu -- CSR k-hop neighbour sampler for B200 (sm_100a): count+scan, row-wise reservoir sampling with the
// reference's exact XORWOW generator assignment, first-occurrence reindex (ordered hash table for the standalone call,
// epoch-tagged direct node map inside a k-hop), the fused k-hop driver, cal_next, and an opt-in O(k) sampler.
//
// What it replaces (reference file:line):
//   TorchQuiver::sample_neighbor / sample_kernel        srcs/cpp/src/quiver/cuda/quiver_sample.cu:113-200
//   CSRRowWiseSampleKernel                              srcs/cpp/include/quiver/cuda_random.cu.hpp:7-69
//   TorchQuiver::reindex_single / reindex_kernel        quiver_sample.cu:305-357, 202-255, FillWithDuplicates :18-63
//   DeviceOrderedHashTable                              srcs/cpp/include/quiver/reindex.cu.hpp:20-158
//   GraphSageSampler.sample's hop loop                  srcs/python/quiver/pyg/sage_sampler.py:118-147
//   cal_next                                            cuda_random.cu.hpp:71-104
//
// Design (B200-first, not a translation):
//   * no per-call cudaMalloc / cudaMemset / thrust: one sampler-owned scratch arena, everything stream-ordered on the
//     caller's stream (the reference uses a private stream pool and ~10 blocking allocations per hop);
//   * every kernel takes its problem size from DEVICE memory, so all hops of a k-hop sample are enqueued back to back
//     and the host synchronises exactly once;
//   * prefix sums are single-pass chained scans (decoupled look-back) fused with the work that produces their input
//     (degree/cap for the sampler, "is first occurrence" for the reindex);
//   * XORWOW states come from a cache (qv_xorwow.cuh) instead of a per-thread skip-ahead every launch;
//   * a lane walks its generator stream through its warp's 16 rows without any warp synchronisation (16 shared-memory
//     reservoirs merged by atomicMax), `r mod m` comes from a fastmod reciprocal table instead of the XU pipe, and the
//     ids to emit are one flat cp.async-staged list per warp;
//   * inside a k-hop the frontier grows incrementally over a persistent node map whose words carry the call's epoch, so
//     nothing is rebuilt or cleared per hop; the dependent kernels are chained with programmatic dependent launch;
//   * int64 ids and 64-bit sizes throughout (the reference truncates to int in several places).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "qv_common.cuh"
#include "qv_xorwow.cuh"

namespace qv
{
namespace
{
// ------------------------------------------------------------------------------------------------------------------
// Device-resident sizes.  meta[kMetaStride*h + ...] for hop h.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMetaS = 0;  // number of seeds of the hop
constexpr int kMetaE = 1;  // number of sampled edges (sum of counts)
constexpr int kMetaF = 2;  // frontier size after reindex
constexpr int kMetaHeavy = 3;  // fused k-hop: rows above kHeavyDeg in the frontier after the hop (sizes the next call's heavy list)
constexpr int kMetaStride = 4;
constexpr int kMetaWords = kMetaStride * (QV_MAX_HOPS + 1);

__device__ __forceinline__ int64_t dev_size(int64_t arg, const int64_t *d)
{
    return d ? *d : arg;
}

// Programmatic dependent launch: the kernels of a k-hop form a strict chain of short launches; launching each with the
// programmatic-serialization attribute lets its blocks be scheduled while the previous kernel drains, and this wait
// (a no-op for ordinary launches) holds them until the predecessor's writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Short kernels additionally release their dependents right away: the next kernel's blocks become resident (parked in
// pdl_wait) while this one still runs, so the hand-over costs no launch latency at all.
__device__ int g_pdl_early = 1;
__device__ __forceinline__ void pdl_release()
{
    if (g_pdl_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    static const bool enabled = !(getenv("QV_PDL") && getenv("QV_PDL")[0] == '0');
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = enabled ? 1 : 0;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// A kernel that synchronises its whole grid (hop_reindex_kernel).  Its grid is bounded by the occupancy, which is enough while
// it is the only such grid on the device; two of them on two streams (two samplers of one process sampling concurrently)
// could each hold half the SMs and spin on the other half forever.  A COOPERATIVE launch makes the driver schedule the
// grid only when every block can be resident at once -- at ~4 us per launch (measured: 3 launches per step, 198-230 ->
// 213-237 us per north-star sample), so it is used when it is needed: `coop` = a second sampler exists on this device
// (or QV_COOP=1; QV_COOP=0 never).  Several PROCESSES sharing a GPU under MPS must set QV_COOP=1.
std::atomic<int> g_live_samplers[64];

template <typename... KArgs, typename... Args>
cudaError_t launch_grid_sync(bool coop, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             Args &&...args)
{
    static const int forced = getenv("QV_COOP") ? atoi(getenv("QV_COOP")) : -1;
    static int mode = 2;  // 2: cooperative + programmatic dependent launch, 1: cooperative only
    if (forced == 0 || (forced < 0 && !coop)) return launch_chained(kernel, grid, block, smem, st, std::forward<Args>(args)...);
    static const bool pdl = !(getenv("QV_PDL") && getenv("QV_PDL")[0] == '0');
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    if (mode == 2) {
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
        if (e == cudaSuccess) return e;
        cudaGetLastError();
        mode = 1;  // this driver does not combine the two attributes: cooperative launches without the early hand-over
    }
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------------------------
// Single-pass chained scan (decoupled look-back).  Tile descriptors: bits 63..62 = flag, 61..0 = value.
// Tiles take their index from an atomic ticket so a tile only ever waits on tiles that are already running.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

// Longest-first scheduling of the sampling kernel.  A row of degree d costs its warp ~(d-k)/32 dependent generator
// draws (parity with the reference's per-lane streams forbids splitting the chain), so the kernel's duration is
// max over warps of (start time + chain): ncu showed one SM still busy for 37 k cycles after all others had finished
// a 106 k-cycle launch.  count_scan therefore lists the warps that own a row with more than kHeavyDraws draws per lane
// (entry = tile * 4 + warp), and the sampling kernel runs those entries in extra blocks at the FRONT of the grid while
// the warp's regular slot finds itself in the list and retires: the long chains start at time zero instead of wherever
// their tile happens to be scheduled.  Same work, same results; list overflow or no list = regular schedule.
constexpr int kHeavyCap = 64;                    // listed warps per launch
constexpr int kHeavyDraws = 48;                  // draws per lane that make a row "heavy" (deg - k > 32 * kHeavyDraws)
constexpr int kHeavyBlocks = kHeavyCap / 4;      // worker blocks: one listed warp per physical warp
// Chain splitting for "mega" rows.  A row with c draws per lane keeps its warp busy for c dependent generator steps (a
// 142 k-degree row: 4460 steps, 61 us alone, ~100 us next to other warps).  The reference's streams cannot be reassigned,
// but XORWOW is linear over GF(2): lane l's state after n draws is A^n * state (xorwow_jump_nib), so the chain is cut into
// segments of kMegaSeg draws whose start states are computed directly; front-of-grid worker warps run the segments, the
// hits meet in a global reservoir through the same commutative atomicMax, and the owner warp jumps its own generators
// over the row and collects the result at write-out time.  Bit-identical to walking the chain.
constexpr int kMegaCap = 8;       // mega rows per launch (more: the rest is walked the ordinary way)
constexpr int kMegaDraws = 1024;  // draws per lane that make a row "mega" (deg - k > 32 * kMegaDraws)
constexpr int kMegaSeg = 256;     // draws per lane per segment
constexpr int kMegaBlocks = 16;   // worker blocks (4 segment-warps each)
constexpr int64_t kMegaMaxDeg = int64_t(1) << 28;  // beyond that the 24-bit jump counter could overflow
// auxiliary words at the end of a scan region (all zeroed with it):
constexpr int kAuxHeavyCount = 0, kAuxHeavyList = 1;                        // [1 .. 1+kHeavyCap)
constexpr int kAuxMegaCount = kAuxHeavyList + kHeavyCap;                     // 65
constexpr int kAuxMegaList = kAuxMegaCount + 1;                              // (row, degree) pairs
constexpr int kAuxMegaDone = kAuxMegaList + 2 * kMegaCap;                    // finished segments per mega row
constexpr int kAuxMegaSlots = kAuxMegaDone + kMegaCap;                       // kMegaCap x 32 uint32 reservoirs
constexpr int kHeavyWords = kAuxMegaSlots + kMegaCap * 16;                   // total

struct ScanState {
    unsigned long long *words;  // [0] = ticket, [1 + tile] = descriptor; zeroed before each launch
    int direct = 0;             // 1: tile = blockIdx.x (the host checked that the whole grid fits on the device at once)
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ long long warp_sum_i64(long long v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Returns the exclusive prefix of this thread's `thread_sum` over the whole grid; `tile` is the ticketed tile index.
// Must be called by all kScanThreads threads.  The last tile stores the grand total to *d_total.
__device__ __forceinline__ long long chained_scan(long long thread_sum, ScanState st, int tile, int n_tiles,
                                                  int64_t *d_total, long long total_base = 0,
                                                  int64_t *d_total_copy = nullptr)
{
    __shared__ long long warp_tot[kScanThreads / 32];
    __shared__ long long tile_excl_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    long long incl = thread_sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();

    long long warp_base = 0, block_agg = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        const long long t = warp_tot[w];
        if (w < warp) warp_base += t;
        block_agg += t;
    }

    if (warp == 0) {
        unsigned long long *desc = st.words + 1;
        long long excl = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile_u64(desc, kFlagPrefix | (static_cast<unsigned long long>(block_agg) & kValueMask));
        } else {
            if (lane == 0)
                st_volatile_u64(desc + tile, kFlagAgg | (static_cast<unsigned long long>(block_agg) & kValueMask));
            long long run = 0;
            int look = tile - 1;
            while (true) {
                const int idx = look - lane;
                unsigned long long w = (idx >= 0) ? ld_volatile_u64(desc + idx) : kFlagPrefix;
                while (__any_sync(0xffffffffu, (w >> 62) == 0)) {
                    if ((w >> 62) == 0) w = ld_volatile_u64(desc + idx);
                }
                const unsigned pm = __ballot_sync(0xffffffffu, (w >> 62) == 2);
                const long long val = static_cast<long long>(w & kValueMask);
                if (pm) {
                    const int first = __ffs(pm) - 1;
                    run += warp_sum_i64(lane <= first ? val : 0);
                    break;
                }
                run += warp_sum_i64(val);
                look -= 32;
            }
            excl = run;
            if (lane == 0)
                st_volatile_u64(desc + tile,
                                kFlagPrefix | (static_cast<unsigned long long>(run + block_agg) & kValueMask));
        }
        if (lane == 0) {
            tile_excl_sh = excl;
            if (tile == n_tiles - 1) {
                if (d_total) *d_total = total_base + excl + block_agg;
                if (d_total_copy) *d_total_copy = total_base + excl + block_agg;
            }
        }
    }
    __syncthreads();
    return tile_excl_sh + warp_base + (incl - thread_sum);
}

// A tile may only wait on tiles that are certain to run.  Tickets guarantee that for any grid size at the cost of one
// atomic round trip before the tile's first load; when every block of the grid can be resident at the same time each
// block is bound to be scheduled whatever the others spin on, so the block index serves as the tile index directly.
__device__ __forceinline__ int take_ticket(ScanState st)
{
    if (st.direct) return static_cast<int>(blockIdx.x);
    __shared__ int tile_sh;
    if (threadIdx.x == 0) tile_sh = static_cast<int>(atomicAdd(st.words, 1ull));
    __syncthreads();
    return tile_sh;
}

// ------------------------------------------------------------------------------------------------------------------
// Epoch-tagged node map of the fused k-hop (one 64-bit word per graph node, never reset).
//   word = (0xFFFFFFFF - epoch) << 32 | payload ; a sample() call uses a fresh epoch, so every word written by an earlier
//   call compares LARGER than anything written now and a plain 64-bit atomicMin overwrites it: no clearing pass, no
//   dependence on the caller's buffers after the call returns.  Within a call:
//     payload = local id            (< 2^31)      the node is in the frontier
//     payload = 2^31 + item index                 candidate: smallest item index that sampled it in this hop
//   so min() keeps a known node known and otherwise elects the first occurrence.  memset(0xFF) = "never seen".
// ------------------------------------------------------------------------------------------------------------------
using MapWord = unsigned long long;
constexpr unsigned int kMapCand = 0x80000000u;
__device__ __forceinline__ MapWord map_word(unsigned int epoch_hi, unsigned int payload)
{
    return (static_cast<MapWord>(epoch_hi) << 32) | payload;
}

// Row r belongs to warp (r & 3) of tile (r >> 6) in the reference geometry (cuda_random.cu.hpp:17-20).
__device__ __forceinline__ void note_heavy_row(unsigned long long *heavy, int64_t r, int64_t deg, int64_t k, int mega_on)
{
    const unsigned long long at = atomicAdd(heavy + kAuxHeavyCount, 1ull);
    if (at < kHeavyCap)
        heavy[kAuxHeavyList + at] = (static_cast<unsigned long long>(r >> 6) << 2) | static_cast<unsigned long long>(r & 3);
    if (mega_on && deg - k > 32 * kMegaDraws && deg < kMegaMaxDeg) {
        const unsigned long long m = atomicAdd(heavy + kAuxMegaCount, 1ull);
        if (m < kMegaCap) {
            heavy[kAuxMegaList + 2 * m] = static_cast<unsigned long long>(r);
            heavy[kAuxMegaList + 2 * m + 1] = static_cast<unsigned long long>(deg);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel A: counts[i] = min(deg(seed_i), k), out_ptr = exclusive scan, total.   (quiver_sample.cu:157-169)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads)
    count_scan_kernel(const int64_t *__restrict__ indptr, int64_t n_nodes, const int64_t *__restrict__ seeds,
                      int64_t S_arg, const int64_t *__restrict__ d_S, int64_t k, int64_t *__restrict__ counts,
                      int64_t *__restrict__ out_ptr, int64_t *__restrict__ d_total, ScanState st, int n_tiles,
                      const int64_t *__restrict__ cached_deg, MapWord *__restrict__ node_map, unsigned int epoch_hi,
                      int64_t *__restrict__ d_err, unsigned long long *__restrict__ heavy, int mega_on)
{
    pdl_wait();
    pdl_release();
    const int64_t S = dev_size(S_arg, d_S);
    const int tile = take_ticket(st);
    const int64_t base = static_cast<int64_t>(tile) * kScanTile + threadIdx.x * kScanItems;
    long long c[kScanItems];
    long long sum = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        const int64_t i = base + j;
        long long v = 0;
        if (i < S) {
            if (cached_deg) {  // hop >= 1 of a fused k-hop: the frontier's degrees were recorded when its nodes joined
                const int64_t deg = cached_deg[i];
                v = (k >= 0 && deg > k) ? k : deg;
                if (heavy && deg - k > 32 * kHeavyDraws) note_heavy_row(heavy, i, deg, k, mega_on);
            } else {
                const int64_t node = seeds[i];
                if (node >= 0 && node < n_nodes) {
                    const int64_t deg = indptr[node + 1] - indptr[node];
                    v = (k >= 0 && deg > k) ? k : deg;
                    if (heavy && deg - k > 32 * kHeavyDraws) note_heavy_row(heavy, i, deg, k, mega_on);
                    if (node_map)  // hop 0: seeds enter the node map
                        atomicMin(&node_map[node], map_word(epoch_hi, kMapCand + static_cast<unsigned int>(i)));
                } else if (node_map) {
                    *d_err = 1;
                }
            }
        }
        c[j] = v;
        sum += v;
    }
    long long excl = chained_scan(sum, st, tile, n_tiles, d_total);
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        const int64_t i = base + j;
        if (i < S) {
            if (counts) counts[i] = c[j];
            out_ptr[i] = excl;
        }
        excl += c[j];
    }
}

// Plain exclusive scan of an int64 array (reindex_single's exclusive_scan(count), quiver_sample.cu:321).
__global__ void __launch_bounds__(kScanThreads)
    plain_scan_kernel(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ out, int64_t *__restrict__ d_total,
                      ScanState st, int n_tiles)
{
    const int tile = take_ticket(st);
    const int64_t base = static_cast<int64_t>(tile) * kScanTile + threadIdx.x * kScanItems;
    long long c[kScanItems];
    long long sum = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        const int64_t i = base + j;
        c[j] = (i < n) ? in[i] : 0;
        sum += c[j];
    }
    long long excl = chained_scan(sum, st, tile, n_tiles, d_total);
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        const int64_t i = base + j;
        if (i < n) out[i] = excl;
        excl += c[j];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel B: row-wise sampling with the reference's generator assignment.
//   virtual block b = 64 consecutive seeds; warp w of the block owns seeds 64b+w, 64b+w+4, ... (<= 16 of them) and its
//   32 lanes own XORWOW streams (seed rand_seed*grid+b, sub-sequence 32w+lane) that persist across those seeds
//   (cuda_random.cu.hpp:17-25,67).  deg <= k: verbatim copy (:33-38).  deg > k: reservoir over positions with
//   max-wins slots (:41-57) then gather (:61-64).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSampleWarps = 4;
constexpr int kSampleTile = 64;
constexpr int kRowsPerWarp = kSampleTile / kSampleWarps;  // 16
constexpr int kSmemSlots = 1024;                          // per warp; larger fan-outs use the output row as slots

// Exact `curand() % m < k` without the XU pipe.  The first profile of this kernel showed the XU pipe 93% busy: nvcc
// lowers a 32-bit `%` by a runtime divisor to I2F + MUFU.RCP + F2I (quarter-rate units).  The divisors a lane meets are
// m = k+1+lane+32j -- a sequence that does not depend on the row -- so the sampler keeps one table
// recip[m] = floor((2^64-1)/m)+1 for every m up to the graph's maximum degree (built once per graph) and uses the
// 64-bit "fastmod" identity (Lemire, Kaser, Kurz 2019): with low = recip[m]*r mod 2^64, r mod m == mulhi64(low, m),
// exact for all 32-bit r, m.  `low < recip[m]*k` is a conservative filter for r mod m < k (never a false negative), so
// the exact remainder is only formed for the rare candidates.  Divisors beyond the table fall back to `%`.
struct RecipTable {
    const unsigned long long *recip;  // [n]; recip[0] = recip[1] = 0
    unsigned int n;
    unsigned int quick;  // 1: short rows take the unrolled path of sample_rows_small_kernel (0 = A-B switch)
    unsigned int limit;  // divisors >= limit bypass the table in sample_rows_small_kernel (experiment switch)
};

// Hint: bring the line holding *p into L1.  A long row walks the reciprocal table linearly (32 consecutive entries per
// draw step); the register ring only covers one trip ahead, which is less than an L2 round trip (mega_probe: 100 cycles per
// draw on a 142 k-degree row instead of 27 while the walk stays inside L1).
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void reservoir_hit(unsigned long long M, uint32_t r, uint32_t m, uint32_t kk, uint32_t idx,
                                              uint32_t *slots)
{
    const unsigned long long low = M * r;
    if (low < M * kk) {
        const uint32_t num = static_cast<uint32_t>(__umul64hi(low, m));
        if (num < kk) atomicMax(&slots[num], idx);
    }
}

// The reservoir loop of one row for one lane: idx = k+lane, k+lane+32, ... < deg; one XORWOW draw per idx, in order.
// The reciprocals are streamed through a 4-deep register ring (loads for draws j+4..j+7 are issued before draws
// j..j+3 are evaluated): a hub row walks the table linearly, and without the ring every iteration of a lone warp
// would stall on an L2 round trip (measured: 3x slower than the XU-bound `%` loop).
template <bool kFast>
__device__ __forceinline__ void reservoir_fill(Xorwow &rng, const RecipTable &rt, uint32_t kk, uint32_t udeg, int lane,
                                               uint32_t *slots)
{
    uint32_t idx = kk + lane;
    if (kFast) {
        const uint32_t fast_end = rt.n > 1 ? min(udeg, rt.n - 1) : 0;  // idx + 1 < rt.n
        if (idx < fast_end) {
            const unsigned long long *tab = rt.recip + 1;  // tab[idx] = recip[idx + 1]
            unsigned long long M0 = tab[idx];
            unsigned long long M1 = idx + 32 < fast_end ? tab[idx + 32] : 0;
            unsigned long long M2 = idx + 64 < fast_end ? tab[idx + 64] : 0;
            unsigned long long M3 = idx + 96 < fast_end ? tab[idx + 96] : 0;
            while (true) {
                const uint32_t nb = idx + 128;
                const unsigned long long N0 = nb < fast_end ? tab[nb] : 0;
                const unsigned long long N1 = nb + 32 < fast_end ? tab[nb + 32] : 0;
                const unsigned long long N2 = nb + 64 < fast_end ? tab[nb + 64] : 0;
                const unsigned long long N3 = nb + 96 < fast_end ? tab[nb + 96] : 0;
                reservoir_hit(M0, xorwow_next(rng), idx + 1, kk, idx, slots);
                if (idx + 32 < fast_end) reservoir_hit(M1, xorwow_next(rng), idx + 33, kk, idx + 32, slots);
                if (idx + 64 < fast_end) reservoir_hit(M2, xorwow_next(rng), idx + 65, kk, idx + 64, slots);
                if (idx + 96 < fast_end) reservoir_hit(M3, xorwow_next(rng), idx + 97, kk, idx + 96, slots);
                if (nb >= fast_end) break;
                idx = nb;
                M0 = N0;
                M1 = N1;
                M2 = N2;
                M3 = N3;
            }
            // first index of this lane's progression at or beyond fast_end
            const uint32_t first = kk + lane;
            idx = first + ((fast_end - first + 31) / 32) * 32;
        }
    }
    for (; idx < udeg; idx += 32) {  // divisors beyond the table (or kFast == false): plain modulo
        const uint32_t num = xorwow_next(rng) % (idx + 1);
        if (num < kk) atomicMax(&slots[num], idx);
    }
}

__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ int64_t row_degree(int64_t r, const int64_t *__restrict__ cached_deg,
                                              const int64_t *__restrict__ seeds, const int64_t *__restrict__ indptr,
                                              int64_t n_nodes)
{
    if (cached_deg) return cached_deg[r];
    const int64_t node = seeds[r];
    return (node >= 0 && node < n_nodes) ? indptr[node + 1] - indptr[node] : 0;
}

// by value in, by value out: taking the generator's address would give it a home in local memory for the whole kernel
__device__ __noinline__ Xorwow xorwow_jump_dev(Xorwow s, uint64_t n, const uint32_t *__restrict__ mats)
{
    xorwow_jump_nib(s, n, mats);
    return s;
}

// Worker side of the chain splitting (see kMegaCap): warp q of the kMegaBlocks front blocks runs segments q, q + W, ...
// of the (row, segment) pairs of the launch's mega rows.  A segment = draws [j * kMegaSeg, (j+1) * kMegaSeg) of every
// lane's share of the row; the lane's generator is positioned by jumping over the draws of the warp's earlier rows
// (computed from their degrees) plus the segment offset.
__device__ __noinline__ void mega_segments(unsigned long long *__restrict__ aux, const uint32_t *__restrict__ rng_states,
                                           const uint32_t *__restrict__ jump_mats, const RecipTable rt, int k,
                                           const int64_t *__restrict__ cached_deg, const int64_t *__restrict__ seeds,
                                           const int64_t *__restrict__ indptr, int64_t n_nodes)
{
    const int lane = threadIdx.x & 31;
    const int n_mega = static_cast<int>(min(aux[kAuxMegaCount], static_cast<unsigned long long>(kMegaCap)));
    if (n_mega == 0) return;
    const uint32_t kk = static_cast<uint32_t>(k);
    unsigned int *slots_g = reinterpret_cast<unsigned int *>(aux + kAuxMegaSlots);
    constexpr int kWorkers = kMegaBlocks * kSampleWarps;
    int pair = static_cast<int>(blockIdx.x) * kSampleWarps + static_cast<int>(threadIdx.x >> 5);
    int base = 0;
    for (int m = 0; m < n_mega; m++) {
        const int64_t r = static_cast<int64_t>(aux[kAuxMegaList + 2 * m]);
        const int64_t deg = static_cast<int64_t>(aux[kAuxMegaList + 2 * m + 1]);
        const uint32_t c0 = static_cast<uint32_t>((deg - kk + 31) >> 5);  // lane 0 draws the most
        const int G = static_cast<int>((c0 + kMegaSeg - 1) / kMegaSeg);
        for (; pair < base + G; pair += kWorkers) {
            const uint32_t j = static_cast<uint32_t>(pair - base);
            const int64_t b = r >> 6;
            const int w = static_cast<int>(r & 3), i = static_cast<int>((r & 63) >> 2);
            int64_t pdeg = 0;
            if (lane < i) pdeg = row_degree(b * kSampleTile + w + static_cast<int64_t>(lane) * kSampleWarps, cached_deg, seeds,
                                            indptr, n_nodes);
            uint64_t n_prev = 0;  // draws this lane's generator made on the warp's earlier rows
            for (int q = 0; q < i; q++) {
                const int64_t dd = min(__shfl_sync(0xffffffffu, pdeg, q), static_cast<int64_t>(0xffffffffu));
                if (dd > static_cast<int64_t>(kk) + lane) n_prev += static_cast<uint64_t>((dd - kk - lane + 31) >> 5);
            }
            const uint32_t c_l = deg > static_cast<int64_t>(kk) + lane ? static_cast<uint32_t>((deg - kk - lane + 31) >> 5) : 0;
            const uint32_t t0 = j * kMegaSeg, t1 = min(t0 + kMegaSeg, c_l);
            if (t0 < t1) {
                Xorwow g;
                const uint32_t *p = rng_states + static_cast<size_t>(b) * kRngStateWords * kRngBlockThreads + (w * 32 + lane);
                g.d = p[0 * kRngBlockThreads];
                g.v0 = p[1 * kRngBlockThreads];
                g.v1 = p[2 * kRngBlockThreads];
                g.v2 = p[3 * kRngBlockThreads];
                g.v3 = p[4 * kRngBlockThreads];
                g.v4 = p[5 * kRngBlockThreads];
                xorwow_jump_nib(g, n_prev + t0, jump_mats);
                unsigned int *srow = slots_g + m * 32;
                uint32_t idx = kk + lane + 32u * t0;
                // plain `%` here: ncu put 43 % of this launch's samples on the wait for the reciprocal-table load (each
                // segment streams through its own cold 64 KB of the table, one load in flight); a handful of warps do
                // not saturate the XU pipe the table was introduced to relieve
                for (uint32_t t = t0; t < t1; t++, idx += 32) {
                    const uint32_t num = xorwow_next(g) % (idx + 1);
                    if (num < kk) atomicMax(&srow[num], idx);
                }
            }
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(aux + kAuxMegaDone + m, 1ull);
        }
        base += G;
    }
}

// Fan-outs up to 32 (every GraphSAGE configuration in BASELINE.json).
//
// Profiling the first two versions of this kernel on the products-shaped bench batch (170 k rows in the last hop) showed
// that the reservoir draws themselves are ~1/4 of the issued instructions (1.8 warp-iterations per row on average); the
// rest was per-row overhead -- shuffles, warp syncs, a dependent memory latency and a 5/32-lane-wide gather per row --
// plus one SM grinding through a 29 k-degree hub row at ~250 cycles per draw long after the others had finished.
// What parity fixes is only this: lane l owns ONE generator stream that serves the lane's draws of row 0, then row 1, ...
// and the number of draws it makes in a row, ceil((deg - k - l) / 32), is known up front.  Hence:
//   * the 16 reservoirs of a warp's rows are all live in shared memory, so the row loop needs NO warp synchronisation:
//     a lane just walks its stream row after row, results meet through shared-memory atomicMax (commutative);
//   * long runs inside one row (hubs) go 8 draws per trip with the fastmod reciprocals of the NEXT trip already in
//     flight and all 8 generator outputs produced before the first test, so a lone warp is not serialised on one
//     divide-and-branch latency per draw; short rows use the plain `%` (nothing to look up);
//   * the ids to emit (verbatim rows and chosen positions alike) are handled as ONE flat list of <= 16*k entries per
//     warp -- 32 lanes wide instead of k lanes wide -- fetched with cp.async into a staging tile (verbatim rows before
//     the generator loop even starts) and written out after a single wait.
template <bool kShortTable, int kHub, int kMinBlocks>
__global__ void __launch_bounds__(kSampleWarps * 32, kMinBlocks)
    sample_rows_small_kernel(const int64_t *__restrict__ indptr, const int64_t *__restrict__ indices, int64_t n_nodes,
                             const int64_t *__restrict__ seeds, int64_t S_arg, const int64_t *__restrict__ d_S, int k,
                             const int64_t *__restrict__ out_ptr, const uint32_t *__restrict__ rng_states,
                             const RecipTable rt, int64_t *__restrict__ out, int64_t *__restrict__ row_out,
                             const int64_t *__restrict__ d_row_off, const int64_t *__restrict__ cached_start,
                             const int64_t *__restrict__ cached_deg, MapWord *__restrict__ node_map, unsigned int epoch_hi,
                             int64_t item_base_arg, const int64_t *__restrict__ d_item_base,
                             int64_t *__restrict__ d_err, unsigned long long *__restrict__ heavy, int late_wait,
                             const uint32_t *__restrict__ jump_mats, int64_t *__restrict__ eid_out,
                             const int64_t *__restrict__ edge_ids)
{
    // dynamic shared memory, sized by the fan-out: per warp 16*k staged ids (8 B), 16*k reservoir slots (4 B) and 16*k
    // entry->row bytes -- 4 KiB per block at k = 5 instead of a fixed 26 KiB, which lifts the occupancy limit
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const uint32_t kcap = k > 0 ? static_cast<uint32_t>(k) : 1u;
    const uint32_t per_warp = kRowsPerWarp * kcap;
    __shared__ int64_t start_sh[kSampleWarps][kRowsPerWarp];
    __shared__ int64_t o_sh[kSampleWarps][kRowsPerWarp];
    __shared__ uint32_t deg_sh[kSampleWarps][kRowsPerWarp];
    __shared__ uint16_t pre_sh[kSampleWarps][kRowsPerWarp + 1];     // entry offset of each row inside the warp's list
    __shared__ int8_t mega_sh[kSampleWarps][kRowsPerWarp];          // index in the launch's mega list, -1 = ordinary row
    // Programmatic dependent launch, taken one step further: the kernel in front of this one is count_scan, whose blocks
    // release their dependents only after they have themselves waited for the hop's inputs (frontier rows, sizes), so
    // those are complete when a block of this kernel starts.  What count_scan PRODUCES (out_ptr, the longest-first list)
    // is needed only for the final write-out -- the wait sits there (late_wait), and count_scan runs underneath this
    // kernel's prologue and generator loop instead of in front of it (QV_PDL_LATE=0 restores the wait at the top).
    if (!late_wait) pdl_wait();
    const int64_t S = dev_size(S_arg, d_S);
    const int lane = threadIdx.x & 31;
    const int wp = threadIdx.x >> 5;  // physical warp: owns a slice of the shared-memory arrays
    int w = wp;                       // logical warp of the reference geometry: decides rows and generator streams
    int64_t b = blockIdx.x;
    bool listed_run = false;
    if (heavy) {
        // front of the grid: [segment workers of the mega rows (only with jump matrices)] [listed heavy warps] [tiles]
        const unsigned int first_listed = jump_mats ? kMegaBlocks : 0;
        if (blockIdx.x < first_listed) {
            pdl_wait();  // the lists are count_scan's output
            mega_segments(heavy, rng_states, jump_mats, rt, k, cached_deg, seeds, indptr, n_nodes);
            return;
        }
        if (blockIdx.x < first_listed + kHeavyBlocks) {  // longest-first schedule
            pdl_wait();
            const unsigned long long n_listed = min(heavy[kAuxHeavyCount], static_cast<unsigned long long>(kHeavyCap));
            const unsigned int slot = (blockIdx.x - first_listed) * kSampleWarps + wp;
            if (slot >= n_listed) return;
            const unsigned long long entry = heavy[kAuxHeavyList + slot];
            b = static_cast<int64_t>(entry >> 2);
            w = static_cast<int>(entry & 3);
            listed_run = true;
        } else {
            b = blockIdx.x - first_listed - kHeavyBlocks;
        }
    }
    if (b * kSampleTile >= S) return;
    const uint32_t kk = static_cast<uint32_t>(k);
    int64_t *stage_w = reinterpret_cast<int64_t *>(dyn_smem) + static_cast<size_t>(wp) * per_warp;
    uint32_t *slots_w = reinterpret_cast<uint32_t *>(dyn_smem + static_cast<size_t>(kSampleWarps) * per_warp * 8) +
                        static_cast<size_t>(wp) * per_warp;
    uint8_t *rowof_w = dyn_smem + static_cast<size_t>(kSampleWarps) * per_warp * 12 + static_cast<size_t>(wp) * per_warp;

    Xorwow rng;
    {
        const uint32_t *p = rng_states + static_cast<size_t>(b) * kRngStateWords * kRngBlockThreads + (w * 32 + lane);
        rng.d = p[0 * kRngBlockThreads];
        rng.v0 = p[1 * kRngBlockThreads];
        rng.v1 = p[2 * kRngBlockThreads];
        rng.v2 = p[3 * kRngBlockThreads];
        rng.v3 = p[4 * kRngBlockThreads];
        rng.v4 = p[5 * kRngBlockThreads];
    }
    // lane i < 16 owns the metadata of row i; one inclusive warp scan lays the rows' entries out back to back
    uint32_t n_entries;
    unsigned int mega_mask = 0;  // bit i: row i is a mega row handled by the segment workers
    {
        int64_t my_start = 0, my_deg = 0, my_o = 0;
        const int64_t r = b * kSampleTile + w + static_cast<int64_t>(lane) * kSampleWarps;
        if (lane < kRowsPerWarp && r < S) {
            if (!late_wait) my_o = out_ptr[r];
            if (cached_deg) {
                my_start = cached_start[r];
                my_deg = cached_deg[r];
            } else {
                const int64_t node = seeds[r];
                if (node >= 0 && node < n_nodes) {
                    my_start = indptr[node];
                    my_deg = indptr[node + 1] - my_start;
                }
            }
        }
        if (heavy && !listed_run && __any_sync(0xffffffffu, my_deg - k > 32 * kHeavyDraws)) {
            // this warp owns a heavy row: if count_scan managed to list it, a front-of-grid block is already on it
            pdl_wait();
            const unsigned long long n_listed = min(heavy[kAuxHeavyCount], static_cast<unsigned long long>(kHeavyCap));
            const unsigned long long me = (static_cast<unsigned long long>(b) << 2) | static_cast<unsigned long long>(w);
            bool found = false;
            for (unsigned int j = lane; j < n_listed; j += 32) found |= heavy[kAuxHeavyList + j] == me;
            if (__any_sync(0xffffffffu, found)) return;
        }
        int my_mega = -1;  // is this lane's row one whose chain the segment workers walk?
        if (jump_mats) {
            const bool cand = my_deg - k > 32 * kMegaDraws && my_deg < kMegaMaxDeg;
            if (__any_sync(0xffffffffu, cand)) {
                pdl_wait();
                const int n_mega = static_cast<int>(min(heavy[kAuxMegaCount], static_cast<unsigned long long>(kMegaCap)));
                if (cand)
                    for (int m = 0; m < n_mega; m++)
                        if (heavy[kAuxMegaList + 2 * m] == static_cast<unsigned long long>(r)) my_mega = m;
            }
        }
        mega_mask = __ballot_sync(0xffffffffu, my_mega >= 0);
        if (lane < kRowsPerWarp) mega_sh[wp][lane] = static_cast<int8_t>(my_mega);
        const uint32_t cnt = static_cast<uint32_t>(my_deg <= k ? my_deg : k);
        uint32_t incl = cnt;
#pragma unroll
        for (int off = 1; off < kRowsPerWarp; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        n_entries = __shfl_sync(0xffffffffu, incl, kRowsPerWarp - 1);
        if (lane < kRowsPerWarp) {
            deg_sh[wp][lane] = static_cast<uint32_t>(min(my_deg, static_cast<int64_t>(0xffffffffu)));
            start_sh[wp][lane] = my_start;
            o_sh[wp][lane] = my_o;
            pre_sh[wp][lane] = static_cast<uint16_t>(incl - cnt);
            for (uint32_t j = 0; j < cnt; j++) rowof_w[incl - cnt + j] = static_cast<uint8_t>(lane);
        }
    }
    {  // every reservoir starts as 0..k-1 (e % kcap by multiply-high: exact for e < 2^16, one division per thread)
        const uint32_t inv_k = kcap > 1 ? 0xFFFFFFFFu / kcap + 1u : 0u;
        for (uint32_t e = lane; e < per_warp; e += 32) slots_w[e] = kcap > 1 ? e - __umulhi(e, inv_k) * kcap : 0u;
    }
    __syncwarp();

    // verbatim rows: their ids can start travelling now
    for (uint32_t e = lane; e < n_entries; e += 32) {
        const int i = rowof_w[e];
        if (deg_sh[wp][i] <= kk) cp_async_8(&stage_w[e], indices + start_sh[wp][i] + (e - pre_sh[wp][i]));
    }

    // this lane's generator stream, row after row, no synchronisation
    {
        const uint32_t first = kk + lane;
        const bool quick_rows = rt.quick != 0;
        const unsigned long long *tab = rt.recip + 1;  // tab[idx] = recip[idx + 1]
        // rt.limit (QV_TAB_LIMIT, default: none) can cut the table short so that long walks fall back to `%`.  Measured:
        // with an 8192-entry limit a 142 k-degree row got 2x SLOWER (540 -> 1150 us in mega_probe with QV_MEGA=0) and the
        // 64 k-seed batch collapsed (8.8 -> 1.0 G SEPS): cold table loads are still cheaper than the divergent `%` loop.
        const uint32_t tab_n = rt.n > 0 ? min(rt.n - 1, rt.limit) : 0;
        for (int i = 0; i < kRowsPerWarp; i++) {
            const uint32_t d = deg_sh[wp][i];
            uint32_t *srow = slots_w + static_cast<size_t>(i) * kcap;

            if (kShortTable && d > kk && d <= tab_n && (d - kk + 31) >> 5 < kHub) {
                // The common row (a few draws per lane): the test above is warp-uniform and the <= kHub-1 draws are fully
                // unrolled and predicated, so there is no divergent loop, no per-lane trip count and no bound check on the
                // table (idx < d <= tab_n); the reciprocal loads do not depend on the generator chain and issue early.
                // (ncu source view of the previous loop: 58 warp instructions per draw round, 14 of them loop control.)
                if (quick_rows) {
#pragma unroll
                    for (int t = 0; t < kHub - 1; t++) {
                        const uint32_t idx = first + 32u * t;
                        if (idx < d) reservoir_hit(tab[idx], xorwow_next(rng), idx + 1, kk, idx, srow);
                    }
                    continue;
                }
            }
            if (d <= first) continue;
            uint32_t rem = (d - first + 31) >> 5, idx = first;
            if (mega_mask && (mega_mask >> i & 1u)) {  // the workers walk this row's chain (every lane has draws): step over
                rng = xorwow_jump_dev(rng, rem, jump_mats);
                continue;
            }
            if (rem >= kHub && idx + 64 * kHub < tab_n) {
                unsigned long long M[kHub];
#pragma unroll
                for (int u = 0; u < kHub; u++) M[u] = tab[idx + 32 * u];
                while (true) {
                    unsigned long long N[kHub];
                    const bool more = rem >= 2 * kHub && idx + 96 * kHub < tab_n;
                    prefetch_l1(tab + min(idx + 32u * kHub * 6u, tab_n - 1));  // six trips ahead (idx + 64 kHub < tab_n here)
                    if (more) {
#pragma unroll
                        for (int u = 0; u < kHub; u++) N[u] = tab[idx + 32 * kHub + 32 * u];
                    }
                    uint32_t r[kHub];
#pragma unroll
                    for (int u = 0; u < kHub; u++) r[u] = xorwow_next(rng);
                    bool cand = false;
                    unsigned long long low[kHub];
#pragma unroll
                    for (int u = 0; u < kHub; u++) {
                        low[u] = M[u] * r[u];
                        cand |= low[u] < M[u] * kk;
                    }
                    if (cand) {
#pragma unroll
                        for (int u = 0; u < kHub; u++) {
                            const uint32_t num = static_cast<uint32_t>(__umul64hi(low[u], idx + 32 * u + 1));
                            if (num < kk) atomicMax(&srow[num], idx + 32 * u);
                        }
                    }
                    idx += 32 * kHub;
                    rem -= kHub;
                    if (!more) break;
#pragma unroll
                    for (int u = 0; u < kHub; u++) M[u] = N[u];
                }
            }
            if (kShortTable) {
                // short rows: reciprocals of small divisors stay L1-resident; loaded one draw ahead
                unsigned long long M = (rem > 0 && idx < tab_n) ? tab[idx] : 0;
                for (; rem > 0; rem--, idx += 32) {
                    const uint32_t r = xorwow_next(rng);
                    const unsigned long long cur = M;
                    const bool in_tab = idx < tab_n;
                    if (rem > 1 && idx + 32 < tab_n) M = tab[idx + 32];
                    if (in_tab) {
                        reservoir_hit(cur, r, idx + 1, kk, idx, srow);
                    } else {
                        const uint32_t num = r % (idx + 1);
                        if (num < kk) atomicMax(&srow[num], idx);
                    }
                }
            } else {
                for (; rem > 0; rem--, idx += 32) {
                    const uint32_t num = xorwow_next(rng) % (idx + 1);
                    if (num < kk) atomicMax(&srow[num], idx);
                }
            }
        }
    }
    __syncwarp();

    if (mega_mask) {  // wait until every segment of this warp's mega rows has been walked
        if (lane < kRowsPerWarp && (mega_mask >> lane & 1u)) {
            const uint32_t c0 = (deg_sh[wp][lane] - kk + 31) >> 5;
            const unsigned long long segments = (c0 + kMegaSeg - 1) / kMegaSeg;
            const unsigned long long *done = heavy + kAuxMegaDone + mega_sh[wp][lane];
            while (ld_volatile_u64(done) < segments) {
            }
        }
        __syncwarp();
    }
    // sampled rows: fetch the chosen positions; then one wait and one coalesced write-out of the whole list
    for (uint32_t e = lane; e < n_entries; e += 32) {
        const int i = rowof_w[e];
        if (deg_sh[wp][i] > kk) {
            const uint32_t j = e - pre_sh[wp][i];
            uint32_t pos = slots_w[static_cast<size_t>(i) * kcap + j];
            if (mega_mask && (mega_mask >> i & 1u))  // global reservoir, zero-initialised: an untouched slot keeps position j
                pos = max(__ldcv(reinterpret_cast<const unsigned int *>(heavy + kAuxMegaSlots) + mega_sh[wp][i] * 32 + j), j);
            cp_async_8(&stage_w[e], indices + start_sh[wp][i] + pos);
        }
    }
    if (late_wait) {  // only now are count_scan's offsets needed
        pdl_wait();
        const int64_t r = b * kSampleTile + w + static_cast<int64_t>(lane) * kSampleWarps;
        if (lane < kRowsPerWarp && r < S) o_sh[wp][lane] = out_ptr[r];
        __syncwarp();
    }
    cp_async_wait_all();
    const int64_t row_off = row_out ? (d_row_off ? *d_row_off : 0) : 0;
    const int64_t item_base = node_map ? (d_item_base ? *d_item_base : item_base_arg) : 0;
    for (uint32_t e = lane; e < n_entries; e += 32) {
        const int i = rowof_w[e];
        const int64_t dst = o_sh[wp][i] + (e - pre_sh[wp][i]);
        const int64_t id = stage_w[e];
        out[dst] = id;
        if (row_out) row_out[row_off + dst] = b * kSampleTile + w + static_cast<int64_t>(i) * kSampleWarps;
        if (eid_out) {  // edge id = CSR position of the pick (the reservoirs are final: recompute it instead of carrying it)
            const uint32_t j = e - pre_sh[wp][i];
            uint32_t pos = j;
            if (deg_sh[wp][i] > kk) {
                pos = slots_w[static_cast<size_t>(i) * kcap + j];
                if (mega_mask && (mega_mask >> i & 1u))
                    pos = max(__ldcv(reinterpret_cast<const unsigned int *>(heavy + kAuxMegaSlots) + mega_sh[wp][i] * 32 + j), j);
            }
            const int64_t p = start_sh[wp][i] + pos;
            eid_out[dst] = edge_ids ? edge_ids[p] : p;
        }
        if (node_map) {  // fused k-hop: the sampled id enters the first-occurrence map right here
            if (static_cast<uint64_t>(id) < static_cast<uint64_t>(n_nodes))
                atomicMin(&node_map[id], map_word(epoch_hi, kMapCand + static_cast<unsigned int>(item_base + dst)));
            else
                *d_err = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Opt-in fast sampling (qv_sampler_set_fast): NOT the reference's random stream.  Same contract -- min(deg, k) distinct
// positions of the row, uniform, verbatim copy when deg <= k -- but position-independent and O(k) per row instead of
// O(deg): the j-th pick of a row is perm(j), where perm is a keyed bijection of [0, deg) (a 4-round Feistel network on
// ceil(log2 deg) bits, cycle-walked into range, keyed by (seed, call counter, node id)).  Distinctness holds by
// construction, there are no generator states and no hub tails (a 142 k-degree row costs what a 40-degree row costs).
// One output entry per thread.  Validated structurally and by a chi-square test, never against the reference's ids.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint64_t feistel_pick(uint64_t j, uint64_t n, uint32_t k0, uint32_t k1)
{
    // bijection on [0, 2^(2h)) with 2^(2h) >= n, then cycle-walk back into [0, n)
    int bits = 64 - __clzll(n - 1);
    if (bits < 2) bits = 2;
    const int h = (bits + 1) >> 1;
    const uint64_t mask = (1ull << h) - 1;
    uint64_t x = j;
    do {
        uint64_t l = x >> h, r = x & mask;
#pragma unroll
        for (int round = 0; round < 4; round++) {
            const uint64_t f = mix32(static_cast<uint32_t>(r) ^ (k0 + 0x9E3779B9u * round)) ^
                               (static_cast<uint64_t>(mix32(static_cast<uint32_t>(r >> 32) ^ k1 ^ round)) << 7);
            const uint64_t nl = r;
            r = (l ^ f) & mask;
            l = nl;
        }
        x = (l << h) | r;
    } while (x >= n);
    return x;
}

__global__ void __launch_bounds__(256)
    sample_rows_fast_kernel(const int64_t *__restrict__ indptr, const int64_t *__restrict__ indices, int64_t n_nodes,
                            const int64_t *__restrict__ seeds, int64_t S_arg, const int64_t *__restrict__ d_S, int64_t k_arg,
                            const int64_t *__restrict__ out_ptr, const int64_t *__restrict__ d_E, uint32_t key0,
                            uint32_t key1, int64_t *__restrict__ out, int64_t *__restrict__ row_out,
                            const int64_t *__restrict__ d_row_off, const int64_t *__restrict__ cached_start,
                            const int64_t *__restrict__ cached_deg, int64_t *__restrict__ eid_out,
                            const int64_t *__restrict__ edge_ids)
{
    pdl_wait();
    const int64_t S = dev_size(S_arg, d_S);
    const int64_t E = *d_E;
    const int64_t k = k_arg < 0 ? INT64_MAX : k_arg;
    const int64_t row_off = row_out ? (d_row_off ? *d_row_off : 0) : 0;
    for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < E;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        // owner row: largest r with out_ptr[r] <= e
        int64_t lo = 0, hi = S;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (out_ptr[mid] <= e)
                lo = mid;
            else
                hi = mid;
        }
        const int64_t r = lo, j = e - out_ptr[r];
        int64_t start, deg, node = 0;
        if (cached_deg) {
            start = cached_start[r];
            deg = cached_deg[r];
            node = seeds[r];
        } else {
            node = seeds[r];
            start = indptr[node];  // rows with entries are in range: count_scan gave out-of-range seeds 0 entries
            deg = indptr[node + 1] - start;
        }
        const int64_t pos = deg <= k ? j
                                     : static_cast<int64_t>(feistel_pick(static_cast<uint64_t>(j), static_cast<uint64_t>(deg),
                                                                         key0 ^ mix32(static_cast<uint32_t>(node)),
                                                                         key1 ^ static_cast<uint32_t>(node >> 32)));
        out[e] = indices[start + pos];
        if (row_out) row_out[row_off + e] = r;
        if (eid_out) eid_out[e] = edge_ids ? edge_ids[start + pos] : start + pos;
    }
}

template <bool kSlotsInSmem, bool kFast = true>
__global__ void __launch_bounds__(kSampleWarps * 32)
    sample_rows_kernel(const int64_t *__restrict__ indptr, const int64_t *__restrict__ indices, int64_t n_nodes,
                       const int64_t *__restrict__ seeds, int64_t S_arg, const int64_t *__restrict__ d_S, int64_t k_arg,
                       const int64_t *__restrict__ out_ptr, const uint32_t *__restrict__ rng_states, const RecipTable rt,
                       int64_t *__restrict__ out, int64_t *__restrict__ row_out, const int64_t *__restrict__ d_row_off,
                       int64_t *__restrict__ eid_out, const int64_t *__restrict__ edge_ids)
{
    __shared__ uint32_t slots_sh[kSlotsInSmem ? kSampleWarps * kSmemSlots : 1];
    const int64_t S = dev_size(S_arg, d_S);
    const int64_t b = blockIdx.x;
    if (b * kSampleTile >= S) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = k_arg < 0 ? INT64_MAX : k_arg;

    Xorwow rng;
    {
        const uint32_t *p = rng_states + static_cast<size_t>(b) * kRngStateWords * kRngBlockThreads + threadIdx.x;
        rng.d = p[0 * kRngBlockThreads];
        rng.v0 = p[1 * kRngBlockThreads];
        rng.v1 = p[2 * kRngBlockThreads];
        rng.v2 = p[3 * kRngBlockThreads];
        rng.v3 = p[4 * kRngBlockThreads];
        rng.v4 = p[5 * kRngBlockThreads];
    }

    // one wave of loads fetches the metadata of all (<= 16) rows this warp owns: lane i holds row i
    int64_t my_start = 0, my_deg = 0, my_o = 0;
    {
        const int64_t r = b * kSampleTile + w + static_cast<int64_t>(lane) * kSampleWarps;
        if (lane < kRowsPerWarp && r < S) {
            const int64_t node = seeds[r];
            my_o = out_ptr[r];
            if (node >= 0 && node < n_nodes) {
                my_start = indptr[node];
                my_deg = indptr[node + 1] - my_start;
            }
        }
    }
    const int64_t row_off = row_out ? (d_row_off ? *d_row_off : 0) : 0;
    uint32_t *slots = slots_sh + (kSlotsInSmem ? w * kSmemSlots : 0);

    for (int i = 0; i < kRowsPerWarp; i++) {
        const int64_t r = b * kSampleTile + w + static_cast<int64_t>(i) * kSampleWarps;
        if (r >= S) break;
        const int64_t start = __shfl_sync(0xffffffffu, my_start, i);
        const int64_t deg = __shfl_sync(0xffffffffu, my_deg, i);
        const int64_t o = __shfl_sync(0xffffffffu, my_o, i);
        const int64_t cnt = deg <= k ? deg : k;
        if (row_out)
            for (int64_t j = lane; j < cnt; j += 32) row_out[row_off + o + j] = r;
        if (deg <= k) {
            for (int64_t j = lane; j < deg; j += 32) {
                out[o + j] = indices[start + j];
                if (eid_out) eid_out[o + j] = edge_ids ? edge_ids[start + j] : start + j;
            }
        } else if (kSlotsInSmem) {
            const uint32_t kk = static_cast<uint32_t>(k);
            for (uint32_t j = lane; j < kk; j += 32) slots[j] = j;
            __syncwarp();
            const uint32_t udeg = static_cast<uint32_t>(deg);
            reservoir_fill<kFast>(rng, rt, kk, udeg, lane, slots);
            __syncwarp();
            for (uint32_t j = lane; j < kk; j += 32) {
                const int64_t p = start + slots[j];
                out[o + j] = indices[p];
                if (eid_out) eid_out[o + j] = edge_ids ? edge_ids[p] : p;
            }
            __syncwarp();
        } else {
            unsigned long long *gs = reinterpret_cast<unsigned long long *>(out + o);
            for (int64_t j = lane; j < k; j += 32) gs[j] = static_cast<unsigned long long>(j);
            __syncwarp();
            for (int64_t idx = k + lane; idx < deg; idx += 32) {
                const uint32_t num = xorwow_next(rng) % static_cast<uint32_t>(idx + 1);
                if (static_cast<int64_t>(num) < k) atomicMax(&gs[num], static_cast<unsigned long long>(idx));
            }
            __syncwarp();
            for (int64_t j = lane; j < k; j += 32) {
                const int64_t pos = static_cast<int64_t>(gs[j]);
                out[o + j] = indices[start + pos];
                if (eid_out) eid_out[o + j] = edge_ids ? edge_ids[start + pos] : start + pos;
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Ordered hash table (first occurrence wins).  Slot = {key, first index, local id}, 16 bytes.
// Capacity is a power of two >= 2*(S+E), derived ON THE DEVICE from the hop's sizes so the fused k-hop path clears
// and probes only what the hop needs.
// ------------------------------------------------------------------------------------------------------------------
struct __align__(16) Slot {
    long long key;
    unsigned int index;
    unsigned int local;
};
constexpr long long kEmptyKey = -1;

__device__ __forceinline__ int table_log2(int64_t n_items, int max_log2)
{
    int lg = 10;
    if (n_items > 512) lg = 64 - __clzll(static_cast<unsigned long long>(2 * n_items - 1));
    return lg > max_log2 ? max_log2 : lg;
}
__device__ __forceinline__ uint64_t table_hash(long long key, int lg)
{
    return (static_cast<uint64_t>(key) * 0x9E3779B97F4A7C15ull) >> (64 - lg);
}

__global__ void __launch_bounds__(256)
    table_clear_kernel(Slot *__restrict__ table, int max_log2, int64_t S_arg, const int64_t *__restrict__ d_S,
                       int64_t E_arg, const int64_t *__restrict__ d_E)
{
    const int64_t n = dev_size(S_arg, d_S) + dev_size(E_arg, d_E);
    const uint64_t cap = 1ull << table_log2(n, max_log2);
    int4 *t = reinterpret_cast<int4 *>(table);
    const int4 empty = make_int4(-1, -1, -1, -1);
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < cap;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        t[i] = empty;
}

// Kernel C: insert concat(seeds, outputs); remember each item's slot.
__global__ void __launch_bounds__(256)
    hash_insert_kernel(const int64_t *__restrict__ seeds, int64_t S_arg, const int64_t *__restrict__ d_S,
                       const int64_t *__restrict__ outputs, int64_t E_arg, const int64_t *__restrict__ d_E,
                       Slot *__restrict__ table, int max_log2, uint32_t *__restrict__ pos)
{
    const int64_t S = dev_size(S_arg, d_S), E = dev_size(E_arg, d_E);
    const int64_t n = S + E;
    const int lg = table_log2(n, max_log2);
    const uint64_t mask = (1ull << lg) - 1;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const long long key = i < S ? seeds[i] : outputs[i - S];
        uint64_t p = table_hash(key, lg);
        while (true) {
            const unsigned long long prev =
                atomicCAS(reinterpret_cast<unsigned long long *>(&table[p].key),
                          static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
            if (prev == static_cast<unsigned long long>(kEmptyKey) || prev == static_cast<unsigned long long>(key)) {
                atomicMin(&table[p].index, static_cast<unsigned int>(i));
                break;
            }
            p = (p + 1) & mask;
        }
        pos[i] = static_cast<uint32_t>(p);
    }
}

// Kernel D: first-occurrence flags -> chained scan -> frontier + local ids.   (quiver_sample.cu:39-61)
__global__ void __launch_bounds__(kScanThreads)
    frontier_scan_kernel(int64_t S_arg, const int64_t *__restrict__ d_S, int64_t E_arg, const int64_t *__restrict__ d_E,
                         Slot *__restrict__ table, const uint32_t *__restrict__ pos, int64_t *__restrict__ frontier,
                         int64_t *__restrict__ d_F, ScanState st, int n_tiles, int64_t *__restrict__ d_next_S)
{
    const int64_t n = dev_size(S_arg, d_S) + dev_size(E_arg, d_E);
    const int tile = take_ticket(st);
    const int64_t base = static_cast<int64_t>(tile) * kScanTile + threadIdx.x * kScanItems;
    long long key[kScanItems];
    uint32_t p[kScanItems];
    bool first[kScanItems];
    long long sum = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        const int64_t i = base + j;
        first[j] = false;
        key[j] = 0;
        p[j] = 0;
        if (i < n) {
            p[j] = pos[i];
            const int4 s = *reinterpret_cast<const int4 *>(&table[p[j]]);
            key[j] = (static_cast<long long>(static_cast<unsigned int>(s.y)) << 32) | static_cast<unsigned int>(s.x);
            first[j] = static_cast<unsigned int>(s.z) == static_cast<unsigned int>(i);
        }
        sum += first[j] ? 1 : 0;
    }
    long long excl = chained_scan(sum, st, tile, n_tiles, d_F, 0, d_next_S);
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        if (first[j]) {
            frontier[excl] = key[j];
            table[p[j]].local = static_cast<unsigned int>(excl);
            excl++;
        }
    }
}

// Kernel E: col_idx[e] = local id of outputs[e]; optionally row_idx[e] by binary search over out_ptr
// (quiver_sample.cu:244-251 and :341-351).
__global__ void __launch_bounds__(256)
    emit_edges_kernel(int64_t S_arg, const int64_t *__restrict__ d_S, int64_t E_arg, const int64_t *__restrict__ d_E,
                      const Slot *__restrict__ table, const uint32_t *__restrict__ pos, int64_t *__restrict__ col_idx,
                      int64_t *__restrict__ row_idx, const int64_t *__restrict__ out_ptr)
{
    const int64_t S = dev_size(S_arg, d_S), E = dev_size(E_arg, d_E);
    for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < E;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        col_idx[e] = table[pos[S + e]].local;
        if (row_idx) {
            // largest i with out_ptr[i] <= e  (seeds with zero count share an offset with their successor)
            int64_t lo = 0, hi = S;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (out_ptr[mid] <= e)
                    lo = mid;
                else
                    hi = mid;
            }
            row_idx[e] = lo;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused k-hop reindex over the epoch-tagged node map (see MapWord above), kept across the hops of a sample() call.
// The frontier of hop h+1 is the frontier of hop h plus the not-yet-seen sampled nodes in first-occurrence order, with
// unchanged local ids for the old ones -- so the hash table rebuilt from scratch every hop (reference:
// quiver_sample.cu:202-255, one cudaMalloc+cudaMemset per call) is unnecessary inside a k-hop sample: per hop only the E
// sampled ids are processed (not S+E), with one 8-byte access each and no probing.  Ids outside [0, n_nodes) raise a flag
// and the call is redone on the hash path (they can only come from invalid user seeds or a corrupt CSR).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    map_insert_kernel(const int64_t *__restrict__ prefix, int64_t P_arg, const int64_t *__restrict__ d_P,
                      const int64_t *__restrict__ outputs, const int64_t *__restrict__ d_E, MapWord *__restrict__ map,
                      unsigned int epoch_hi, int64_t n_nodes, int64_t *__restrict__ d_err)
{
    pdl_wait();
    pdl_release();
    const int64_t P = prefix ? dev_size(P_arg, d_P) : 0, E = *d_E;
    const int64_t n = P + E;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t key = i < P ? prefix[i] : outputs[i - P];
        if (static_cast<uint64_t>(key) >= static_cast<uint64_t>(n_nodes)) {
            *d_err = 1;
            continue;
        }
        atomicMin(&map[key], map_word(epoch_hi, kMapCand + static_cast<unsigned int>(i)));
    }
}

// kItems ids per thread: 4 for small hops, 16 for very large ones (fewer tiles in the look-back chain).
// kRows: also record the CSR row (start, degree) of every node that joins, for the next hop's count / sample kernels.
template <int kItems, bool kRows>
__global__ void __launch_bounds__(kScanThreads, kRows ? 4 : 6)
    map_scan_kernel(const int64_t *__restrict__ prefix, int64_t P_arg, const int64_t *__restrict__ d_P,
                    const int64_t *__restrict__ outputs, const int64_t *__restrict__ d_E, MapWord *__restrict__ map,
                    unsigned int epoch_hi, int64_t n_nodes, const int64_t *__restrict__ d_F_prev,
                    int64_t *__restrict__ frontier, int64_t *__restrict__ d_F, ScanState st, int n_tiles,
                    const int64_t *__restrict__ indptr, int64_t *__restrict__ fr_start, int64_t *__restrict__ fr_deg,
                    int64_t *__restrict__ d_next_S)
{
    pdl_wait();
    pdl_release();
    const int64_t P = prefix ? dev_size(P_arg, d_P) : 0, E = *d_E;
    const int64_t n = P + E;
    const long long F_prev = d_F_prev ? *d_F_prev : 0;
    const int tile = take_ticket(st);
    const int64_t base = static_cast<int64_t>(tile) * (kScanThreads * kItems) + threadIdx.x * kItems;
    long long key[kItems];
    bool first[kItems];
    long long sum = 0;
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        const int64_t i = base + j;
        first[j] = false;
        key[j] = 0;
        if (i < n) {
            key[j] = i < P ? prefix[i] : outputs[i - P];
            if (static_cast<uint64_t>(key[j]) < static_cast<uint64_t>(n_nodes))
                first[j] = map[key[j]] == map_word(epoch_hi, kMapCand + static_cast<unsigned int>(i));
        }
        sum += first[j] ? 1 : 0;
    }
    // a node that joins the frontier gets its CSR row located now (the loads overlap the look-back), so the next hop's
    // count / sample kernels read (start, degree) with coalesced loads instead of two dependent random ones per seed
    long long rs[kItems], rd[kItems];
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        rs[j] = rd[j] = 0;
        if (kRows && first[j]) {
            rs[j] = indptr[key[j]];
            rd[j] = indptr[key[j] + 1] - rs[j];
        }
    }
    long long local = F_prev + chained_scan(sum, st, tile, n_tiles, d_F, F_prev, d_next_S);
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        if (first[j]) {
            frontier[local] = key[j];
            map[key[j]] = map_word(epoch_hi, static_cast<unsigned int>(local));
            if (kRows) {
                fr_start[local] = rs[j];
                fr_deg[local] = rd[j];
            }
            local++;
        }
    }
}

__global__ void __launch_bounds__(256)
    map_emit_kernel(const int64_t *__restrict__ outputs, const int64_t *__restrict__ d_E, const MapWord *__restrict__ map,
                    int64_t n_nodes, int64_t *__restrict__ col_idx)
{
    pdl_wait();
    pdl_release();
    const int64_t E = *d_E;
    for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < E;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t key = outputs[e];
        col_idx[e] = static_cast<uint64_t>(key) < static_cast<uint64_t>(n_nodes)
                         ? static_cast<int64_t>(static_cast<unsigned int>(map[key]) & 0x7FFFFFFFu)
                         : 0;
    }
}

// Every hop's emit in one launch at the end of the k-hop.  A node keeps the local id it got when it was first seen
// (later hops only atomicMin candidates > 2^31 over it), so the col_idx of hop h can be read from the map any time after
// hop h's scan: nothing on the hop-to-hop critical path waits for it.  Thread t serves edge t - begin[h] of hop h, with
// begin[] the prefix of the hops' static edge bounds.
struct EmitAll {
    const int64_t *nbr[QV_MAX_HOPS];
    const int64_t *d_E[QV_MAX_HOPS];
    int64_t *col[QV_MAX_HOPS];
    int64_t begin[QV_MAX_HOPS + 1];
    int n_hops;
};

__global__ void __launch_bounds__(256)
    map_emit_all_kernel(const __grid_constant__ EmitAll p, const MapWord *__restrict__ map, int64_t n_nodes)
{
    pdl_wait();
    pdl_release();
    const int64_t total = p.begin[p.n_hops];
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int h = 0;
#pragma unroll
        for (int i = 1; i < QV_MAX_HOPS; i++)
            if (i < p.n_hops && t >= p.begin[i]) h = i;
        const int64_t e = t - p.begin[h];
        if (e >= *p.d_E[h]) continue;
        const int64_t key = p.nbr[h][e];
        p.col[h][e] = static_cast<uint64_t>(key) < static_cast<uint64_t>(n_nodes)
                          ? static_cast<int64_t>(static_cast<unsigned int>(map[key]) & 0x7FFFFFFFu)
                          : 0;
    }
}

#include "qv_hop_kernels.cuh"

__global__ void __launch_bounds__(256)
    max_degree_kernel(const int64_t *__restrict__ indptr, int64_t n_nodes, unsigned long long *__restrict__ result)
{
    long long best = 0;
    for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < n_nodes;
         v += static_cast<int64_t>(gridDim.x) * blockDim.x)
        best = max(best, static_cast<long long>(indptr[v + 1] - indptr[v]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(result, static_cast<unsigned long long>(best));
}

__global__ void __launch_bounds__(256) recip_table_kernel(unsigned long long *__restrict__ recip, unsigned int n)
{
    for (unsigned int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x)
        recip[m] = m < 2 ? 0ull : (0xFFFFFFFFFFFFFFFFull / m + 1ull);
}

// ------------------------------------------------------------------------------------------------------------------
// cal_next (cuda_random.cu.hpp:71-104): one hop of access-probability propagation,
//   p'[v] = 1 - (1 - p[v]) * prod_{u in N(v)} skip(u),   skip(u) = 1 - p[u] * min(1, k / deg u)   (1 when deg u = 0).
// The reference runs one THREAD per node with a serial loop over its neighbours (three dependent random reads per
// neighbour, a 142 k-neighbour hub in a single thread).  Here a warp owns a node: 32 neighbours' skip factors are
// fetched in parallel, then multiplied IN CSR ORDER (each lane replays the same 32 sequential multiplies through
// shuffles; idle lanes contribute exactly 1.0f), so the fp32 result is bit-identical to the serial loop.  Arithmetic is
// spelled with round-to-nearest intrinsics; the final 1 - a*b is the single FMA nvcc contracts in the reference kernel.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    cal_next_kernel(const float *__restrict__ last_prob, float *__restrict__ cur_prob, int64_t N, int k,
                    const int64_t *__restrict__ indptr, const int64_t *__restrict__ indices)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp; row < N; row += n_warps) {
        const int64_t start = indptr[row];
        const int64_t deg = indptr[row + 1] - start;
        if (deg == 0) {
            if (lane == 0) cur_prob[row] = 0;
            continue;
        }
        float acc = 1.0f;
        for (int64_t base = 0; base < deg; base += 32) {
            float skip = 1.0f;
            if (base + lane < deg) {
                const int64_t u = indices[start + base + lane];
                const int64_t ustart = indptr[u];
                const int64_t udeg = indptr[u + 1] - ustart;
                if (udeg != 0) {
                    const float p = last_prob[u];
                    if (udeg <= k)
                        skip = __fsub_rn(1.0f, p);
                    else
                        skip = __fadd_rn(__fsub_rn(1.0f, p),
                                         __fdiv_rn(__fmul_rn(p, static_cast<float>(udeg - k)), static_cast<float>(udeg)));
                }
            }
#pragma unroll
            for (int l = 0; l < 32; l++) acc = __fmul_rn(acc, __shfl_sync(0xffffffffu, skip, l));
        }
        if (lane == 0) cur_prob[row] = __fmaf_rn(-__fsub_rn(1.0f, last_prob[row]), acc, 1.0f);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------------
struct Buffer {
    void *ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return QV_OK;
        size_t want = cap ? cap : 4096;
        while (want < bytes) want *= 2;
        if (ptr) {
            QV_CUDA(cudaFree(ptr));
            ptr = nullptr;
            cap = 0;
        }
        QV_CUDA(cudaMalloc(&ptr, want));
        cap = want;
        return QV_OK;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

// Blocks of `kernel` (kScanThreads threads, no dynamic shared memory) that can be resident on the device at once.
template <typename K>
int resident_capacity(K kernel, int n_sm)
{
    static int per_sm = -1;  // one instance per kernel type
    if (per_sm < 0) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kScanThreads, 0) != cudaSuccess) {
            cudaGetLastError();
            v = 0;
        }
        per_sm = v;
    }
    static const bool off = getenv("QV_SCAN_TICKETS") && getenv("QV_SCAN_TICKETS")[0] == '1';  // A-B switch
    return off ? 0 : per_sm * n_sm;
}

inline int tiles_for(int64_t n) { return static_cast<int>(std::max<int64_t>(1, (n + kScanTile - 1) / kScanTile)); }
inline int host_log2_cap(int64_t n_items)
{
    int lg = 10;
    if (n_items > 512) lg = 64 - __builtin_clzll(static_cast<unsigned long long>(2 * n_items - 1));
    return lg;
}
}  // namespace
}  // namespace qv

using namespace qv;

struct qv_sampler {
    int device = 0;
    const int64_t *indptr = nullptr;
    const int64_t *indices = nullptr;
    int64_t n_nodes = 0, n_edges = 0;
    int n_sm = 148;

    int64_t *d_meta = nullptr;  // kMetaWords device scalars
    int64_t *h_meta = nullptr;  // pinned mirror
    cudaEvent_t meta_ready = nullptr;
    Buffer scan;                // two scan-state regions
    size_t scan_region_words = 0;
    Buffer table;  // Slot[2^table_log2]
    int table_log2 = 0;
    Buffer pos;      // uint32[n items]
    Buffer out_ptr;  // int64[S]      (fused path / reindex_single)
    Buffer nbr;      // int64[E]      (fused path: sampled neighbour ids)
    Buffer rng_mats;   // XORWOW skip matrices (device copy)
    Buffer jump_mats;  // XORWOW jump tables (A^(2^i) as 4-bit lookups; uploaded when a graph has mega rows)
    Buffer rng_cache;  // states for rand_seed == 0, blocks [0, rng_cache_blocks)
    int64_t rng_cache_blocks = 0;
    Buffer rng_tmp;  // states for rand_seed != 0 (per launch)
    int heavy_front = kHeavyFrontInit;  // heavy blocks in front of a fused hop's sampling grid (kHeavyFrontInit, grown on demand)
    Buffer ctl;       // control words of the two-kernels-per-hop path (heavy list, tile descriptors, barrier counters)
    Buffer tgt;       // int32[E]: target row of every sampled edge of the current hop (two-kernels-per-hop path)
    Buffer tile_base; // int64[tiles]: output offset of every 64-row tile of the next hop (written by hop_reindex_kernel)
    Buffer fr_meta;   // [2][bound] int64: CSR row start / degree of every frontier node (fused k-hop path)
    Buffer node_map;  // MapWord per graph node: epoch-tagged first-occurrence map of the fused k-hop path
    unsigned int map_epoch = 0;  // 0 = the map has never been initialised
    bool err_dirty = false;      // the device error flag (kMetaErr) is set and must be cleared before the next k-hop
    bool fast = false;           // opt-in non-reference sampling (qv_sampler_set_fast)
    uint64_t fast_calls = 0;     // call counter mixed into the fast sampler's key
    Buffer recip;    // fastmod reciprocals for divisors [0, recip_n)
    unsigned int recip_n = 0;
    int64_t max_degree = 0;
    const int64_t *edge_ids = nullptr;  // optional user edge ids per CSR position (qv_sampler_set_edge_ids); borrowed
    bool counted = false;               // this object is in g_live_samplers (see launch_grid_sync)
};

namespace
{
constexpr int kScanRegions = 2 * QV_MAX_HOPS;  // two scans per hop, every hop of a fused k-hop has its own pair

int ensure_scan(qv_sampler *s, int64_t max_items)
{
    const size_t words = static_cast<size_t>(tiles_for(max_items)) + 2 + kHeavyWords;  // descriptors + longest-first list
    const size_t region = (words + 15) & ~size_t(15);
    if (region > s->scan_region_words) {
        QV_TRY(s->scan.ensure(region * kScanRegions * sizeof(unsigned long long)));
        s->scan_region_words = s->scan.cap / (kScanRegions * sizeof(unsigned long long));
    }
    return QV_OK;
}
ScanState scan_region(qv_sampler *s, int which)
{
    return ScanState{static_cast<unsigned long long *>(s->scan.ptr) + which * s->scan_region_words};
}
unsigned long long *heavy_region(qv_sampler *s, int which)  // the last kHeavyWords of a scan region
{
    return scan_region(s, which).words + s->scan_region_words - kHeavyWords;
}
int zero_scan_regions(qv_sampler *s, int64_t items0, int64_t items1, cudaStream_t st)
{
    // only the descriptors that will be used need zeroing
    const size_t w0 = static_cast<size_t>(tiles_for(items0)) + 2, w1 = static_cast<size_t>(tiles_for(items1)) + 2;
    QV_CUDA(cudaMemsetAsync(scan_region(s, 0).words, 0, w0 * 8, st));
    if (items1 >= 0) QV_CUDA(cudaMemsetAsync(scan_region(s, 1).words, 0, w1 * 8, st));
    return QV_OK;
}
int ensure_table(qv_sampler *s, int64_t max_items)
{
    const int lg = host_log2_cap(max_items);
    if (lg > s->table_log2) {
        QV_REQUIRE(lg <= 33, "reindex: %lld items exceed the hash table limit", (long long)max_items);
        QV_TRY(s->table.ensure((size_t(1) << lg) * sizeof(Slot)));
        s->table_log2 = lg;
    }
    QV_TRY(s->pos.ensure(static_cast<size_t>(std::max<int64_t>(max_items, 1)) * sizeof(uint32_t)));
    return QV_OK;
}

// XORWOW states for a launch over (up to) `rows_bound` seeds.  rand_seed == 0 (the reference's literal) is served
// from a cache that only ever grows; other seeds are generated per launch.
int rng_states_for(qv_sampler *s, uint64_t rand_seed, int64_t rows_arg, const int64_t *d_rows, int64_t rows_bound,
                   cudaStream_t st, const uint32_t **states)
{
    const int64_t blocks = (rows_bound + kSampleTile - 1) / kSampleTile;
    const size_t bytes_per_block = size_t(kRngStateWords) * kRngBlockThreads * sizeof(uint32_t);
    if (!s->rng_mats.ptr) {
        const size_t bytes = size_t(kXorwowBits) * kXorwowWords * kRngBlockThreads * sizeof(uint32_t);
        QV_TRY(s->rng_mats.ensure(bytes));
        QV_CUDA(cudaMemcpy(s->rng_mats.ptr, xorwow_subseq_matrices_host(), bytes, cudaMemcpyHostToDevice));
    }
    if (rand_seed == 0) {
        if (blocks > s->rng_cache_blocks) {
            int64_t want = std::max<int64_t>(blocks, 1024);
            want = std::max(want, s->rng_cache_blocks * 2);
            // growing re-allocates: make sure nothing in flight still reads the old cache
            QV_CUDA(cudaStreamSynchronize(st));
            QV_TRY(s->rng_cache.ensure(static_cast<size_t>(want) * bytes_per_block));
            QV_TRY(xorwow_fill_states(static_cast<const uint32_t *>(s->rng_mats.ptr), 0, want * kSampleTile, nullptr,
                                      want, static_cast<uint32_t *>(s->rng_cache.ptr), st));
            s->rng_cache_blocks = want;
        }
        *states = static_cast<const uint32_t *>(s->rng_cache.ptr);
        return QV_OK;
    }
    if (static_cast<size_t>(blocks) * bytes_per_block > s->rng_tmp.cap) QV_CUDA(cudaStreamSynchronize(st));
    QV_TRY(s->rng_tmp.ensure(static_cast<size_t>(std::max<int64_t>(blocks, 1)) * bytes_per_block));
    QV_TRY(xorwow_fill_states(static_cast<const uint32_t *>(s->rng_mats.ptr), rand_seed, rows_arg, d_rows, blocks,
                              static_cast<uint32_t *>(s->rng_tmp.ptr), st));
    *states = static_cast<const uint32_t *>(s->rng_tmp.ptr);
    return QV_OK;
}

inline unsigned grid_for(int64_t items, int threads, int n_sm, int waves = 8)
{
    const int64_t blocks = (items + threads - 1) / threads;
    return static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(blocks, int64_t(n_sm) * waves)));
}

struct HopExtras {  // fused k-hop only; all null for the standalone calls
    const int64_t *cached_start = nullptr, *cached_deg = nullptr;
    MapWord *node_map = nullptr;
    unsigned int epoch_hi = 0;
    int64_t item_base = 0;
    const int64_t *d_item_base = nullptr;
    int64_t *d_err = nullptr;
    unsigned long long *heavy = nullptr;  // longest-first list (kHeavyWords, zeroed): filled by count_scan, read by the sampler
    int late_wait = 0;                    // the sampling kernel waits for count_scan only before its write-out
    const uint32_t *jump_mats = nullptr;  // XORWOW jump matrices: non-null enables chain splitting of mega rows
    int64_t *eid_out = nullptr;           // optional edge-id output of the hop (any caller, not only the fused k-hop)
};

int launch_count_scan(qv_sampler *s, const int64_t *seeds, int64_t S_arg, const int64_t *d_S, int64_t S_bound,
                      int64_t k, int64_t *counts, int64_t *out_ptr, int64_t *d_total, int region, cudaStream_t st,
                      const HopExtras &x = HopExtras())
{
    const int n_tiles = tiles_for(S_bound);
    ScanState scan = scan_region(s, region);
    scan.direct = n_tiles <= resident_capacity(count_scan_kernel, s->n_sm);
    QV_CUDA(launch_chained(count_scan_kernel, n_tiles, kScanThreads, 0, st, s->indptr, s->n_nodes, seeds, S_arg, d_S, k,
                           counts, out_ptr, d_total, scan, n_tiles, x.cached_deg,
                           x.cached_deg ? nullptr : x.node_map, x.epoch_hi, x.d_err, x.heavy, x.jump_mats ? 1 : 0));
    QV_CHECK_LAUNCH("count_scan_kernel");
    return QV_OK;
}

int launch_sample(qv_sampler *s, const int64_t *seeds, int64_t S_arg, const int64_t *d_S, int64_t S_bound, int64_t k,
                  uint64_t rand_seed, const int64_t *out_ptr, int64_t *out, int64_t *row_out, const int64_t *d_row_off,
                  cudaStream_t st, const HopExtras &x = HopExtras(), bool *fused_insert = nullptr,
                  const int64_t *d_E = nullptr, int64_t E_bound = 0)
{
    if (fused_insert) *fused_insert = false;
    if (S_bound <= 0) return QV_OK;
    if (s->fast && d_E) {
        if (E_bound <= 0) return QV_OK;
        const uint64_t c = s->fast_calls++;
        const uint64_t h = (rand_seed ^ (c * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull;
        QV_CUDA(launch_chained(sample_rows_fast_kernel, grid_for(E_bound, 256, s->n_sm), 256, 0, st, s->indptr, s->indices,
                               s->n_nodes, seeds, S_arg, d_S, k, out_ptr, d_E, static_cast<uint32_t>(h),
                               static_cast<uint32_t>(h >> 32), out, row_out, d_row_off, x.cached_start, x.cached_deg,
                               x.eid_out, s->edge_ids));
        QV_CHECK_LAUNCH("sample_rows_fast_kernel");
        return QV_OK;
    }
    const uint32_t *states = nullptr;
    QV_TRY(rng_states_for(s, rand_seed, S_arg, d_S, S_bound, st, &states));
    const int64_t blocks = (S_bound + kSampleTile - 1) / kSampleTile;
    QV_REQUIRE(blocks < (int64_t(1) << 31), "sample: too many seeds (%lld)", (long long)S_bound);
    static const bool quick_off = getenv("QV_SAMPLE_QUICK") && getenv("QV_SAMPLE_QUICK")[0] == '0';
    static const unsigned int tab_limit = getenv("QV_TAB_LIMIT") ? static_cast<unsigned int>(atoll(getenv("QV_TAB_LIMIT")))
                                                                    : 0xFFFFFFFFu;  // no limit
    const RecipTable rt{static_cast<const unsigned long long *>(s->recip.ptr), s->recip_n, quick_off ? 0u : 1u, tab_limit};
    static const int impl = getenv("QV_SAMPLE_IMPL") ? atoi(getenv("QV_SAMPLE_IMPL")) : 0;  // tuning switch
    const size_t small_smem = static_cast<size_t>(kSampleWarps) * kRowsPerWarp * std::max<int64_t>(k, 1) * 13;
    if (k >= 0 && k <= 32 && !(impl & 1)) {
        if (!(impl & 4))  // default: fastmod table for short rows too (measured -10 us per bench step vs plain %)
            QV_CUDA(launch_chained(sample_rows_small_kernel<true, 4, 8>,
                                   static_cast<unsigned>(blocks + (x.heavy ? kHeavyBlocks : 0) + (x.jump_mats ? kMegaBlocks : 0)),
                                   kSampleWarps * 32,
                                   small_smem, st, s->indptr, s->indices, s->n_nodes, seeds, S_arg, d_S,
                                   static_cast<int>(k), out_ptr, states, rt, out, row_out, d_row_off, x.cached_start,
                                   x.cached_deg, x.node_map, x.epoch_hi, x.item_base, x.d_item_base, x.d_err, x.heavy,
                                   x.late_wait, x.jump_mats, x.eid_out, s->edge_ids));
        else
            sample_rows_small_kernel<false, 4, 8><<<static_cast<unsigned>(blocks + (x.heavy ? kHeavyBlocks : 0) +
                                                                          (x.jump_mats ? kMegaBlocks : 0)),
                                                    kSampleWarps * 32, small_smem, st>>>(
                s->indptr, s->indices, s->n_nodes, seeds, S_arg, d_S, static_cast<int>(k), out_ptr, states, rt, out,
                row_out, d_row_off, x.cached_start, x.cached_deg, x.node_map, x.epoch_hi, x.item_base, x.d_item_base, x.d_err,
                x.heavy, 0, x.jump_mats, x.eid_out, s->edge_ids);
        if (fused_insert) *fused_insert = x.node_map != nullptr && x.d_err != nullptr;
    } else if (impl & 2) {
        sample_rows_kernel<true, false><<<static_cast<unsigned>(blocks), kSampleWarps * 32, 0, st>>>(
            s->indptr, s->indices, s->n_nodes, seeds, S_arg, d_S, k, out_ptr, states, rt, out, row_out, d_row_off, x.eid_out,
            s->edge_ids);
    } else if (k < 0 || k <= kSmemSlots) {
        sample_rows_kernel<true><<<static_cast<unsigned>(blocks), kSampleWarps * 32, 0, st>>>(
            s->indptr, s->indices, s->n_nodes, seeds, S_arg, d_S, k, out_ptr, states, rt, out, row_out, d_row_off, x.eid_out,
            s->edge_ids);
    } else {
        sample_rows_kernel<false><<<static_cast<unsigned>(blocks), kSampleWarps * 32, 0, st>>>(
            s->indptr, s->indices, s->n_nodes, seeds, S_arg, d_S, k, out_ptr, states, rt, out, row_out, d_row_off, x.eid_out,
            s->edge_ids);
    }
    QV_CHECK_LAUNCH("sample_rows_kernel");
    return QV_OK;
}


// clear + insert + frontier scan + emit, sizes from device scalars (or host args when the pointers are null)
int launch_reindex(qv_sampler *s, const int64_t *seeds, int64_t S_arg, const int64_t *d_S, int64_t S_bound,
                   const int64_t *outputs, int64_t E_arg, const int64_t *d_E, int64_t E_bound, int64_t *frontier,
                   int64_t *d_F, int64_t *col_idx, int64_t *row_idx, const int64_t *out_ptr, int scan_which,
                   cudaStream_t st, int64_t *d_next_S = nullptr)
{
    const int64_t n_bound = S_bound + E_bound;
    Slot *table = static_cast<Slot *>(s->table.ptr);
    uint32_t *pos = static_cast<uint32_t *>(s->pos.ptr);
    const uint64_t cap_bound = 1ull << host_log2_cap(n_bound);
    table_clear_kernel<<<grid_for(static_cast<int64_t>(cap_bound), 256, s->n_sm), 256, 0, st>>>(
        table, s->table_log2, S_arg, d_S, E_arg, d_E);
    QV_CHECK_LAUNCH("table_clear_kernel");
    if (n_bound > 0) {
        hash_insert_kernel<<<grid_for(n_bound, 256, s->n_sm), 256, 0, st>>>(seeds, S_arg, d_S, outputs, E_arg, d_E,
                                                                              table, s->table_log2, pos);
        QV_CHECK_LAUNCH("hash_insert_kernel");
    }
    const int n_tiles = tiles_for(n_bound);
    frontier_scan_kernel<<<n_tiles, kScanThreads, 0, st>>>(S_arg, d_S, E_arg, d_E, table, pos, frontier, d_F,
                                                            scan_region(s, scan_which), n_tiles, d_next_S);
    QV_CHECK_LAUNCH("frontier_scan_kernel");
    if (E_bound > 0) {
        emit_edges_kernel<<<grid_for(E_bound, 256, s->n_sm), 256, 0, st>>>(S_arg, d_S, E_arg, d_E, table, pos, col_idx,
                                                                            row_idx, out_ptr);
        QV_CHECK_LAUNCH("emit_edges_kernel");
    }
    return QV_OK;
}
}  // namespace

extern "C" {

int qv_sampler_create(int device, const int64_t *indptr, int64_t n_nodes, const int64_t *indices, int64_t n_edges,
                      qv_sampler **out)
{
    QV_REQUIRE(out != nullptr, "qv_sampler_create: out is NULL");
    *out = nullptr;
    QV_REQUIRE(indptr != nullptr && n_nodes >= 0 && n_edges >= 0, "qv_sampler_create: bad CSR arguments");
    QV_REQUIRE(indices != nullptr || n_edges == 0, "qv_sampler_create: indices is NULL");
    DeviceGuard g(device);
    qv_sampler *s = new (std::nothrow) qv_sampler();
    QV_REQUIRE(s != nullptr, "qv_sampler_create: out of host memory");
    s->device = device;
    s->indptr = indptr;
    s->indices = indices;
    s->n_nodes = n_nodes;
    s->n_edges = n_edges;
    s->n_sm = sm_count(device);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&s->d_meta), kMetaWords * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemset(s->d_meta, 0, kMetaWords * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&s->h_meta), kMetaWords * sizeof(int64_t), 0);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->meta_ready, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (s->d_meta) cudaFree(s->d_meta);
        delete s;
        return fail(QV_ERR_CUDA, "qv_sampler_create: %s", cudaGetErrorString(e));
    }
    // largest degree of the graph -> size of the divisor table used by the reservoir loop (one pass over indptr)
    {
        unsigned long long *d_max = reinterpret_cast<unsigned long long *>(s->d_meta);
        if (n_nodes > 0) {
            max_degree_kernel<<<grid_for(n_nodes, 256, s->n_sm, 8), 256>>>(indptr, n_nodes, d_max);
            count_launch();
        }
        unsigned long long h_max = 0;
        e = cudaMemcpy(&h_max, d_max, sizeof h_max, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemset(s->d_meta, 0, sizeof(int64_t));
        if (e == cudaSuccess) {
            s->max_degree = static_cast<int64_t>(h_max);
            const unsigned long long want = std::min<unsigned long long>(h_max + 2, 1ull << 24);
            int rc = s->recip.ensure(static_cast<size_t>(want) * sizeof(unsigned long long));
            if (rc != QV_OK) {
                qv_sampler_destroy(s);
                return rc;
            }
            s->recip_n = static_cast<unsigned int>(want);
            recip_table_kernel<<<grid_for(static_cast<int64_t>(want), 256, s->n_sm, 8), 256>>>(
                static_cast<unsigned long long *>(s->recip.ptr), s->recip_n);
            count_launch();
            e = cudaDeviceSynchronize();
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            qv_sampler_destroy(s);
            return fail(QV_ERR_CUDA, "qv_sampler_create: %s", cudaGetErrorString(e));
        }
    }
    {
        const char *env = getenv("QV_PDL_EARLY");
        const int early = (env && env[0] == '0') ? 0 : 1;
        cudaMemcpyToSymbol(g_pdl_early, &early, sizeof early);
        const int hop_dbg = getenv("QV_HOP_DEBUG") ? atoi(getenv("QV_HOP_DEBUG")) : 0;
        cudaMemcpyToSymbol(g_hop_debug, &hop_dbg, sizeof hop_dbg);
    }
    if (device >= 0 && device < 64) g_live_samplers[device].fetch_add(1);
    s->counted = true;
    *out = s;
    return QV_OK;
}

int qv_sampler_destroy(qv_sampler *s)
{
    if (!s) return QV_OK;
    if (s->counted && s->device >= 0 && s->device < 64) g_live_samplers[s->device].fetch_sub(1);
    DeviceGuard g(s->device);
    cudaDeviceSynchronize();
    s->scan.release();
    s->table.release();
    s->pos.release();
    s->out_ptr.release();
    s->nbr.release();
    s->rng_mats.release();
    s->jump_mats.release();
    s->rng_cache.release();
    s->rng_tmp.release();
    s->recip.release();
    s->node_map.release();
    s->fr_meta.release();
    s->ctl.release();
    s->tgt.release();
    s->tile_base.release();
    if (s->d_meta) cudaFree(s->d_meta);
    if (s->h_meta) cudaFreeHost(s->h_meta);
    if (s->meta_ready) cudaEventDestroy(s->meta_ready);
    delete s;
    return QV_OK;
}

int qv_sample_count(qv_sampler *s, const int64_t *seeds, int64_t S, int64_t k, int64_t *counts, int64_t *out_ptr,
                    int64_t *total, qv_stream_t stream)
{
    QV_REQUIRE(s && total, "qv_sample_count: NULL argument");
    QV_REQUIRE(S >= 0, "qv_sample_count: negative seed count");
    *total = 0;
    if (S == 0) return QV_OK;  // the reference launches a 0-block grid here (quiver.cu.hpp:388)
    QV_REQUIRE(seeds && counts && out_ptr, "qv_sample_count: NULL array");
    DeviceGuard g(s->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QV_TRY(ensure_scan(s, S));
    QV_TRY(zero_scan_regions(s, S, -1, st));
    QV_TRY(launch_count_scan(s, seeds, S, nullptr, S, k, counts, out_ptr, s->d_meta + kMetaE, 0, st));
    QV_CUDA(cudaMemcpyAsync(s->h_meta + kMetaE, s->d_meta + kMetaE, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    QV_CUDA(cudaStreamSynchronize(st));
    *total = s->h_meta[kMetaE];
    return QV_OK;
}

int qv_sampler_set_edge_ids(qv_sampler *s, const int64_t *edge_ids)
{
    QV_REQUIRE(s != nullptr, "qv_sampler_set_edge_ids: NULL sampler");
    s->edge_ids = edge_ids;
    return QV_OK;
}

int qv_sample_fill(qv_sampler *s, const int64_t *seeds, int64_t S, int64_t k, uint64_t rand_seed,
                   const int64_t *out_ptr, int64_t *neighbors, int64_t *edge_ids_out, qv_stream_t stream)
{
    QV_REQUIRE(s, "qv_sample_fill: NULL sampler");
    if (S <= 0) return QV_OK;
    QV_REQUIRE(seeds && out_ptr, "qv_sample_fill: NULL array");
    QV_REQUIRE(k < (int64_t(1) << 31), "qv_sample_fill: fan-out %lld too large", (long long)k);
    DeviceGuard g(s->device);
    // fast mode needs the total (left in d_meta by the qv_sample_count call that sized `neighbors`)
    HopExtras x;
    x.eid_out = edge_ids_out;
    return launch_sample(s, seeds, S, nullptr, S, k, rand_seed, out_ptr, neighbors, nullptr, nullptr,
                         static_cast<cudaStream_t>(stream), x, nullptr, s->d_meta + kMetaE,
                         s->fast ? s->h_meta[kMetaE] : 0);
}

int qv_reindex(qv_sampler *s, const int64_t *inputs, int64_t S, const int64_t *outputs, int64_t tot,
               const int64_t *counts, int64_t *frontier, int64_t *row_idx, int64_t *col_idx, int64_t *n_frontier,
               qv_stream_t stream)
{
    QV_REQUIRE(s && n_frontier, "qv_reindex: NULL argument");
    QV_REQUIRE(S >= 0 && tot >= 0, "qv_reindex: negative size");
    *n_frontier = 0;
    if (S + tot == 0) return QV_OK;
    QV_REQUIRE(S + tot < (int64_t(1) << 32) - 1, "qv_reindex: %lld items exceed 2^32", (long long)(S + tot));
    QV_REQUIRE(frontier && (inputs || S == 0) && (outputs || tot == 0), "qv_reindex: NULL array");
    QV_REQUIRE((row_idx && col_idx && counts) || tot == 0, "qv_reindex: NULL edge arrays");
    DeviceGuard g(s->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QV_TRY(ensure_scan(s, S + tot));
    QV_TRY(ensure_table(s, S + tot));
    QV_TRY(s->out_ptr.ensure(static_cast<size_t>(std::max<int64_t>(S, 1)) * sizeof(int64_t)));
    QV_TRY(zero_scan_regions(s, S, S + tot, st));
    int64_t *optr = static_cast<int64_t *>(s->out_ptr.ptr);
    if (S > 0 && tot > 0) {
        const int n_tiles = tiles_for(S);
        plain_scan_kernel<<<n_tiles, kScanThreads, 0, st>>>(counts, S, optr, nullptr, scan_region(s, 0), n_tiles);
        QV_CHECK_LAUNCH("plain_scan_kernel");
    }
    QV_TRY(launch_reindex(s, inputs, S, nullptr, S, outputs, tot, nullptr, tot, frontier, s->d_meta + kMetaF, col_idx,
                          row_idx, optr, 1, st));
    QV_CUDA(cudaMemcpyAsync(s->h_meta + kMetaF, s->d_meta + kMetaF, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    QV_CUDA(cudaStreamSynchronize(st));
    *n_frontier = s->h_meta[kMetaF];
    return QV_OK;
}

int qv_khop_bounds(int64_t S, const int64_t *sizes, int n_hops, int64_t *bound_nodes, int64_t *bound_edges)
{
    QV_REQUIRE(sizes && bound_nodes && bound_edges, "qv_khop_bounds: NULL argument");
    QV_REQUIRE(n_hops >= 1 && n_hops <= QV_MAX_HOPS, "qv_khop_bounds: n_hops must be in [1, %d]", QV_MAX_HOPS);
    QV_REQUIRE(S >= 0, "qv_khop_bounds: negative seed count");
    bound_nodes[0] = S;
    for (int h = 0; h < n_hops; h++) {
        if (sizes[h] < 0)
            return fail(QV_ERR_UNSUPPORTED, "qv_khop: sizes[%d] = %lld has no static bound; use the per-hop calls", h,
                        (long long)sizes[h]);
        const __int128 e = static_cast<__int128>(bound_nodes[h]) * sizes[h];
        if (e + bound_nodes[h] >= (static_cast<__int128>(1) << 32) - 1)
            return fail(QV_ERR_UNSUPPORTED, "qv_khop: hop %d bound exceeds 2^32 items; use the per-hop calls", h);
        bound_edges[h] = static_cast<int64_t>(e);
        bound_nodes[h + 1] = bound_nodes[h] + bound_edges[h];
    }
    return QV_OK;
}

namespace
{
constexpr int kMetaErr = kMetaStride * QV_MAX_HOPS + 3;  // a free device scalar: "id outside [0, n_nodes) seen"

// Feature gather to enqueue behind the last hop (qv_khop_gather); table == nullptr: none.
struct GatherTail {
    const qv_shard_table *table = nullptr;
    const int64_t *feature_order = nullptr;
    int64_t row_bytes = 0;
    void *features = nullptr;
    int64_t capacity_rows = 0;  // rows `features` can hold (<= the static frontier bound)
    int variant = 0;
};

// One attempt of the fused k-hop.  use_map: direct node map (default) or the per-hop hash table.
int khop_run(qv_sampler *s, const int64_t *seeds, int64_t S, const int64_t *sizes, int n_hops, uint64_t rand_seed,
             int64_t *n_id, int64_t *const *edge_buf, int64_t *const *eid_buf, const int64_t *bn, const int64_t *be,
             bool use_map, cudaStream_t st, bool *id_error, const GatherTail &tail)
{
    int64_t *optr = static_cast<int64_t *>(s->out_ptr.ptr);
    int64_t *nbr_base = static_cast<int64_t *>(s->nbr.ptr);
    MapWord *map = static_cast<MapWord *>(s->node_map.ptr);
    static const bool emit_per_hop = getenv("QV_EMIT_PER_HOP") && getenv("QV_EMIT_PER_HOP")[0] == '1';  // A-B switch
    EmitAll emit;
    memset(&emit, 0, sizeof emit);
    emit.n_hops = n_hops;
    int64_t *d_err = s->d_meta + kMetaErr;
    *id_error = false;
    unsigned int epoch_hi = 0;
    if (use_map) {
        if (s->map_epoch == 0 || s->map_epoch >= 0xFFFFFFF0u) {  // first use, or the 32-bit epoch is about to wrap
            QV_CUDA(cudaMemsetAsync(map, 0xFF, static_cast<size_t>(std::max<int64_t>(s->n_nodes, 1)) * sizeof(MapWord), st));
            s->map_epoch = 0;
        }
        s->map_epoch++;  // every call (also a failed one) gets its own epoch: older words never need clearing
        epoch_hi = 0xFFFFFFFFu - s->map_epoch;
    }
    if (s->err_dirty) {  // the flag is only ever raised by a call that saw a bad id: clear it after such a call, not always
        QV_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int64_t), st));
        s->err_dirty = false;
    }
    // one memset clears the scan descriptors of every hop (each hop owns regions 2h and 2h+1)
    if (s->scan_region_words * 2 * n_hops <= (size_t(1) << 20)) {
        QV_CUDA(cudaMemsetAsync(s->scan.ptr, 0, s->scan_region_words * 2 * n_hops * sizeof(unsigned long long), st));
    } else {
        for (int h = 0; h < n_hops; h++) {
            QV_CUDA(cudaMemsetAsync(scan_region(s, 2 * h).words, 0, (tiles_for(bn[h]) + 2) * sizeof(unsigned long long), st));
            QV_CUDA(cudaMemsetAsync(heavy_region(s, 2 * h), 0, kHeavyWords * sizeof(unsigned long long), st));
            QV_CUDA(cudaMemsetAsync(scan_region(s, 2 * h + 1).words, 0,
                                    (tiles_for(bn[h] + be[h]) + 2) * sizeof(unsigned long long), st));
        }
    }
    for (int h = 0; h < n_hops; h++) {
        int64_t *m = s->d_meta + kMetaStride * h;
        // hop 0's seed count is known on the host; later hops read the previous hop's frontier size on the device
        const int64_t S_h = h == 0 ? S : 0;
        const int64_t *d_S = h == 0 ? nullptr : m + kMetaS;
        int64_t *d_E = m + kMetaE;
        int64_t *d_F = m + kMetaF;
        const int64_t *hop_seeds = h == 0 ? seeds : n_id;
        int64_t *d_next_S = m + kMetaStride + kMetaS;  // the next hop's seed count is this hop's frontier size
        // the sampled ids of every hop stay alive until the single emit at the end (map path); the hash path reuses one region
        int64_t *nbr = nbr_base + (use_map ? emit.begin[h] : 0);
        emit.nbr[h] = nbr;
        emit.d_E[h] = d_E;
        emit.col[h] = edge_buf[h];
        emit.begin[h + 1] = emit.begin[h] + be[h];
        HopExtras x;
        x.eid_out = eid_buf ? eid_buf[h] : nullptr;
        int64_t *fr_start = static_cast<int64_t *>(s->fr_meta.ptr), *fr_deg = fr_start ? fr_start + bn[n_hops] : nullptr;
        if (use_map) {
            x.node_map = map;
            x.epoch_hi = epoch_hi;
            x.d_err = d_err;
            x.item_base = h == 0 ? S : 0;  // items of hop 0 are [seeds | outputs]; later hops: outputs only
            if (h >= 1 && fr_start) {
                x.cached_start = fr_start;
                x.cached_deg = fr_deg;
            }
        }
        // longest-first schedule of the sampling kernel (small-fan-out kernel only; QV_HEAVY_FIRST=0 is the A-B switch)
        static const bool heavy_off = getenv("QV_HEAVY_FIRST") && getenv("QV_HEAVY_FIRST")[0] == '0';
        static const int impl_sw = getenv("QV_SAMPLE_IMPL") ? atoi(getenv("QV_SAMPLE_IMPL")) : 0;
        if (!heavy_off && !s->fast && sizes[h] >= 0 && sizes[h] <= 32 && !(impl_sw & 1) &&
            s->max_degree - sizes[h] > 32 * kHeavyDraws)
            x.heavy = heavy_region(s, 2 * h);
        // chain splitting of mega rows (needs the jump matrices on the device; QV_MEGA=0 is the A-B switch)
        static const bool mega_off = getenv("QV_MEGA") && getenv("QV_MEGA")[0] == '0';
        if (x.heavy && !mega_off && s->max_degree - sizes[h] > 32 * kMegaDraws) {
            if (!s->jump_mats.ptr) {
                const size_t bytes = kJumpTableWords * sizeof(uint32_t);  // the 4-bit lookup form of A^(2^i)
                QV_TRY(s->jump_mats.ensure(bytes));
                QV_CUDA(cudaMemcpy(s->jump_mats.ptr, xorwow_jump_tables_host(), bytes, cudaMemcpyHostToDevice));
            }
            x.jump_mats = static_cast<const uint32_t *>(s->jump_mats.ptr);
        }
        // count_scan directly in front of the sampling kernel (rand_seed 0: no state-fill kernel in between): overlap them
        static const bool late_off = getenv("QV_PDL_LATE") && getenv("QV_PDL_LATE")[0] == '0';
        x.late_wait = (!late_off && rand_seed == 0 && !s->fast) ? 1 : 0;
        bool fused_insert = false;
        QV_TRY(launch_count_scan(s, hop_seeds, S_h, d_S, bn[h], sizes[h], nullptr, optr, d_E, 2 * h, st, x));
        // Inserting the sampled ids into the node map from inside the sampling kernel: for a large hop it was measured
        // slower (+14 us on the kernel's critical blocks vs 9 us for a separate, perfectly parallel insert kernel), for a
        // small hop it was neutral (saved launch vs longer kernel): off unless QV_FUSE_INSERT_BELOW is set.
        static const int64_t fuse_below = getenv("QV_FUSE_INSERT_BELOW") ? atoll(getenv("QV_FUSE_INSERT_BELOW")) : 0;
        HopExtras xs = x;
        if (bn[h] > fuse_below) xs.node_map = nullptr;
        // edge_buf[h] = [col (source local ids) | row (target = seed position)], each E long, E read on the device
        QV_TRY(launch_sample(s, hop_seeds, S_h, d_S, bn[h], sizes[h], rand_seed, optr, nbr, edge_buf[h], d_E, st, xs,
                             &fused_insert, d_E, be[h]));
        if (!use_map) {
            QV_TRY(launch_reindex(s, hop_seeds, S_h, d_S, bn[h], nbr, 0, d_E, be[h], n_id, d_F, edge_buf[h], nullptr,
                                  nullptr, 2 * h + 1, st, d_next_S));
        } else {
            // hop 0 also enters the seeds (they become local ids 0..S-1, duplicates merged); later hops only add
            const int64_t *prefix = h == 0 ? seeds : nullptr;
            const int64_t items = (h == 0 ? bn[0] : 0) + be[h];
            if (items > 0 && !fused_insert) {  // fan-outs > 32 use the generic sampling kernel, which does not insert
                QV_CUDA(launch_chained(map_insert_kernel, grid_for(items, 256, s->n_sm), 256, 0, st, prefix, S_h, d_S, nbr,
                                       d_E, map, epoch_hi, s->n_nodes, d_err));
                QV_CHECK_LAUNCH("map_insert_kernel");
            }
            int64_t *fs = h + 1 < n_hops ? fr_start : nullptr, *fd = h + 1 < n_hops ? fr_deg : nullptr;
            const bool big = items > (int64_t(4) << 20);  // below that, more (smaller) tiles hide the random-load latency better
            const int per_tile = kScanThreads * (big ? 16 : 4);
            const int n_tiles = static_cast<int>(std::max<int64_t>(1, (items + per_tile - 1) / per_tile));
            ScanState scan = scan_region(s, 2 * h + 1);
#define QV_MAP_SCAN(ITEMS, ROWS)                                                                                         \
    do {                                                                                                                 \
        scan.direct = n_tiles <= resident_capacity(map_scan_kernel<ITEMS, ROWS>, s->n_sm);                              \
        QV_CUDA(launch_chained(map_scan_kernel<ITEMS, ROWS>, n_tiles, kScanThreads, 0, st, prefix, S_h, d_S, nbr, d_E,   \
                               map, epoch_hi, s->n_nodes, h == 0 ? nullptr : d_S, n_id, d_F, scan, n_tiles, s->indptr, \
                               fs, fd, d_next_S));                                                                      \
    } while (0)
            if (big && fd) QV_MAP_SCAN(16, true);
            else if (big) QV_MAP_SCAN(16, false);
            else if (fd) QV_MAP_SCAN(4, true);
            else QV_MAP_SCAN(4, false);
#undef QV_MAP_SCAN
            QV_CHECK_LAUNCH("map_scan_kernel");
            if (emit_per_hop && be[h] > 0) {
                QV_CUDA(launch_chained(map_emit_kernel, grid_for(be[h], 256, s->n_sm), 256, 0, st, nbr, d_E, map,
                                       s->n_nodes, edge_buf[h]));
                QV_CHECK_LAUNCH("map_emit_kernel");
            }
        }
    }
    if (use_map && !emit_per_hop && emit.begin[n_hops] > 0) {
        QV_CUDA(launch_chained(map_emit_all_kernel, grid_for(emit.begin[n_hops], 256, s->n_sm), 256, 0, st, emit, map,
                               s->n_nodes));
        QV_CHECK_LAUNCH("map_emit_all_kernel");
    }
    QV_CUDA(cudaMemcpyAsync(s->h_meta, s->d_meta, kMetaWords * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    QV_CUDA(cudaEventRecord(s->meta_ready, st));
    if (tail.table) {
        // The sizes are already on their way to the host; the gather goes in behind them with the frontier size read on
        // the device (grid sized for the static bound), so the host wakes up -- and the caller builds its tensors --
        // while the rows are being copied.  After an id error n_id holds arbitrary values: the gather range-checks every
        // id (zero rows), and the whole call is redone by the caller.
        const int64_t *d_n = s->d_meta + kMetaStride * (n_hops - 1) + kMetaF;
        QV_TRY(gather_enqueue(tail.table, n_id, tail.feature_order, std::min(bn[n_hops], tail.capacity_rows), d_n,
                              tail.row_bytes, tail.features, tail.variant, st));
    }
    QV_CUDA(cudaEventSynchronize(s->meta_ready));
    if (use_map && s->h_meta[kMetaErr] != 0) {
        *id_error = true;  // the caller redoes the call on the hash path
        s->err_dirty = true;
    }
    return QV_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Two kernels per hop (qv_hop_kernels.cuh): hop_sample_kernel + hop_reindex_kernel.  Applies when every fan-out is in
// [0, 32], the generator is the reference's (not the opt-in fast sampler) and every hop's item bound fits one
// co-resident reindex grid; anything else takes khop_run (round 1's chain of single-purpose kernels).
// ------------------------------------------------------------------------------------------------------------------
int blocks_per_sm(const void *kernel, int threads, size_t smem)
{
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        v = 0;
    }
    return v;
}

struct ReindexPlan {
    int items = 0;  // items per thread (template argument); 0 = does not fit
    int grid = 0;
};

ReindexPlan plan_reindex(qv_sampler *s, int64_t items_bound)
{
    // grid: one 512-thread block per 512 items of the BOUND, at most two blocks per SM (all resident: the kernel uses grid
    // barriers); items per thread: what the bound then needs, rounded up to an instantiated size
    static int per_sm = -1;
    if (per_sm < 0) per_sm = blocks_per_sm(reinterpret_cast<const void *>(hop_reindex_kernel<16>), kReindexThreads, 0);
    ReindexPlan p;
    const int64_t cap = std::min<int64_t>(int64_t(std::min(per_sm, 2)) * s->n_sm, kReindexMaxBlocks);
    if (cap <= 0) return p;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(cap, (items_bound + kReindexThreads - 1) / kReindexThreads));
    const int64_t need = (((items_bound + grid - 1) / grid) + kReindexThreads - 1) / kReindexThreads;
    for (int c : {1, 2, 4, 8, 16})
        if (need <= c) {
            p.items = c;
            p.grid = static_cast<int>(grid);
            return p;
        }
    return p;
}

bool fused_hops_apply(qv_sampler *s, int64_t S, const int64_t *sizes, int n_hops, const int64_t *bn, const int64_t *be)
{
    static const bool off = getenv("QV_KHOP_FUSED") && getenv("QV_KHOP_FUSED")[0] == '0';  // A-B switch
    if (off || s->fast) return false;
    for (int h = 0; h < n_hops; h++) {
        if (sizes[h] < 0 || sizes[h] > 32) return false;
        const ReindexPlan plan = plan_reindex(s, be[h] + (h == 0 ? S : 0));
        if (plan.items == 0) return false;
        if (h + 1 < n_hops) {  // the reindex kernel of hop h scans the next hop's tiles, at most kReindexTilesPerBlock per block
            const int64_t tiles_next = (bn[h + 1] + kSampleTile - 1) / kSampleTile;
            if ((tiles_next + plan.grid - 1) / plan.grid > kReindexTilesPerBlock) return false;
        }
        if ((bn[h] + kSampleTile - 1) / kSampleTile >= (int64_t(1) << 30)) return false;
    }
    return true;
}

int khop_run_fused(qv_sampler *s, const int64_t *seeds, int64_t S, const int64_t *sizes, int n_hops, uint64_t rand_seed,
                   int64_t *n_id, int64_t *const *edge_buf, int64_t *const *eid_buf, const int64_t *bn, const int64_t *be,
                   cudaStream_t st, bool *id_error, const GatherTail &tail)
{
    int64_t *nbr = static_cast<int64_t *>(s->nbr.ptr);
    int32_t *tgt = static_cast<int32_t *>(s->tgt.ptr);
    MapWord *map = static_cast<MapWord *>(s->node_map.ptr);
    int64_t *d_err = s->d_meta + kMetaErr;
    *id_error = false;
    if (s->map_epoch == 0 || s->map_epoch >= 0xFFFFFFF0u) {  // first use, or the 32-bit epoch is about to wrap
        QV_CUDA(cudaMemsetAsync(map, 0xFF, static_cast<size_t>(std::max<int64_t>(s->n_nodes, 1)) * sizeof(MapWord), st));
        s->map_epoch = 0;
    }
    s->map_epoch++;
    const unsigned int epoch_hi = 0xFFFFFFFFu - s->map_epoch;
    if (s->err_dirty) {
        QV_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int64_t), st));
        s->err_dirty = false;
    }
    // control words: [heavy list | per hop: ticket, barrier counter, block counts, tile descriptors]
    int64_t max_tiles = 1;
    for (int h = 0; h < n_hops; h++) max_tiles = std::max(max_tiles, (bn[h] + kSampleTile - 1) / kSampleTile);
    const size_t stride = (static_cast<size_t>(kHopCtlFixed + max_tiles) + 7) & ~size_t(7);
    const size_t ctl_words = kCtlHeader + stride * n_hops;
    unsigned long long *ctl = static_cast<unsigned long long *>(s->ctl.ptr);
    QV_CUDA(cudaMemsetAsync(ctl, 0, ctl_words * sizeof(unsigned long long), st));

    int64_t *fr_start = static_cast<int64_t *>(s->fr_meta.ptr), *fr_deg = fr_start + bn[n_hops];
    // 35 KB of static + 4-53 KB of dynamic shared memory per 256-thread block: ask for the largest shared-memory carve-out
    // (with the default one ncu showed the shared-memory occupancy limit at half the register limit)
    static const bool three = getenv("QV_HOP_MINB") && getenv("QV_HOP_MINB")[0] == '3';  // A-B switch: 3 resident blocks per SM
    void (*const sample_kernel)(HopSampleArgs) = three ? hop_sample_kernel<4, 3> : hop_sample_kernel<4, 4>;
    static std::atomic<unsigned long long> attr_set{0};  // per device (function attributes belong to the device's copy of the kernel)
    if (s->device < 0 || s->device >= 64 || !(attr_set.load() >> s->device & 1ull)) {
        cudaFuncSetAttribute(sample_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        // 35 KB static (tiles + ring) + up to 53 KB dynamic at k = 32 exceeds the 48 KB a kernel gets without opting in
        QV_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        if (s->device >= 0 && s->device < 64) attr_set.fetch_or(1ull << s->device);
    }
    static const int sample_per_sm = blocks_per_sm(reinterpret_cast<const void *>(sample_kernel), kHopThreads, 0);
    for (int h = 0; h < n_hops; h++) {
        int64_t *m = s->d_meta + kMetaStride * h;
        unsigned long long *hop = ctl + kCtlHeader + stride * h;
        const int k = static_cast<int>(sizes[h]);
        const uint32_t *states = nullptr;
        QV_TRY(rng_states_for(s, rand_seed, h == 0 ? S : 0, h == 0 ? nullptr : m + kMetaS, bn[h], st, &states));
        HopSampleArgs a;
        memset(&a, 0, sizeof a);
        a.indptr = s->indptr;
        a.indices = s->indices;
        a.n_nodes = s->n_nodes;
        a.seeds = h == 0 ? seeds : n_id;
        a.S_arg = h == 0 ? S : 0;
        a.d_S = h == 0 ? nullptr : m + kMetaS;
        a.k = k;
        a.rng_states = states;
        a.rt = RecipTable{static_cast<const unsigned long long *>(s->recip.ptr), s->recip_n, 1u, 0xFFFFFFFFu};
        a.out = nbr;
        a.tgt = tgt;
        a.eid_out = eid_buf ? eid_buf[h] : nullptr;
        a.edge_ids = s->edge_ids;
        a.cached_start = h == 0 ? nullptr : fr_start;
        a.cached_deg = h == 0 ? nullptr : fr_deg;
        a.desc = hop + kHopCtlFixed;
        a.heavy = ctl;
        a.d_E = m + kMetaE;
        a.tile_base = h == 0 ? nullptr : static_cast<const int64_t *>(s->tile_base.ptr);
        static const bool insert_in_reindex = getenv("QV_HOP_INSERT") && getenv("QV_HOP_INSERT")[0] == 'r';  // A-B switch
        if (!insert_in_reindex) {
            a.node_map = map;
            a.epoch_hi = epoch_hi;
            a.item_base = h == 0 ? S : 0;
            a.d_err = d_err;
        }
        a.n_front = (h == 0 || s->max_degree <= kHeavyDeg) ? 0 : s->heavy_front;
        const int64_t tiles = (bn[h] + kSampleTile - 1) / kSampleTile;
        const int64_t blocks = (tiles + kHopTiles - 1) / kHopTiles;
        const size_t smem = static_cast<size_t>(kHopWarps) * kRowsPerWarp * std::max(k, 1) * 13;
        a.ticket = (h == 0 && blocks > int64_t(sample_per_sm) * s->n_sm) ? hop : nullptr;  // only hop 0 looks back
        // a grid of at most ~3 blocks per SM leaves registers and threads for the reindex kernel's two 512-thread blocks
        static const bool early_off = getenv("QV_HOP_EARLY") && getenv("QV_HOP_EARLY")[0] == '0';
        a.release_early = (!early_off && blocks + a.n_front <= int64_t(3) * s->n_sm / 2) ? 1 : 0;
        QV_CUDA(launch_chained(sample_kernel, static_cast<unsigned>(blocks + a.n_front), kHopThreads, smem, st, a));
        QV_CHECK_LAUNCH("hop_sample_kernel");

        HopReindexArgs r;
        memset(&r, 0, sizeof r);
        r.prefix = h == 0 ? seeds : nullptr;
        r.P_arg = h == 0 ? S : 0;
        r.nbr = nbr;
        r.d_E = m + kMetaE;
        r.map = map;
        r.epoch_hi = epoch_hi;
        r.insert_done = insert_in_reindex ? 0 : 1;
        r.n_nodes = s->n_nodes;
        r.d_F_prev = h == 0 ? nullptr : m + kMetaS;
        r.frontier = n_id;
        r.d_F = m + kMetaF;
        r.d_next_S = m + kMetaStride + kMetaS;
        r.indptr = s->indptr;
        r.fr_start = h + 1 < n_hops ? fr_start : nullptr;
        r.fr_deg = h + 1 < n_hops ? fr_deg : nullptr;
        r.heavy = ctl;
        r.heavy_cap = static_cast<unsigned int>(s->heavy_front);
        r.d_heavy_seen = m + kMetaHeavy;
        r.tgt = tgt;
        r.edge_buf = edge_buf[h];
        r.bar = hop + 1;
        r.agg = hop + 2;
        r.d_err = d_err;
        if (h + 1 < n_hops) {
            r.tile_base_next = static_cast<int64_t *>(s->tile_base.ptr);
            r.d_E_next = m + kMetaStride + kMetaE;
            r.k_next = sizes[h + 1];
        }
        const ReindexPlan plan = plan_reindex(s, be[h] + (h == 0 ? S : 0));
        const bool coop = s->device >= 0 && s->device < 64 && g_live_samplers[s->device].load() > 1;
        switch (plan.items) {
        case 1: QV_CUDA(launch_grid_sync(coop, hop_reindex_kernel<1>, plan.grid, kReindexThreads, 0, st, r)); break;
        case 2: QV_CUDA(launch_grid_sync(coop, hop_reindex_kernel<2>, plan.grid, kReindexThreads, 0, st, r)); break;
        case 4: QV_CUDA(launch_grid_sync(coop, hop_reindex_kernel<4>, plan.grid, kReindexThreads, 0, st, r)); break;
        case 8: QV_CUDA(launch_grid_sync(coop, hop_reindex_kernel<8>, plan.grid, kReindexThreads, 0, st, r)); break;
        default: QV_CUDA(launch_grid_sync(coop, hop_reindex_kernel<16>, plan.grid, kReindexThreads, 0, st, r)); break;
        }
        QV_CHECK_LAUNCH("hop_reindex_kernel");
    }
    QV_CUDA(cudaMemcpyAsync(s->h_meta, s->d_meta, kMetaWords * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    QV_CUDA(cudaEventRecord(s->meta_ready, st));
    if (tail.table) {
        const int64_t *d_n = s->d_meta + kMetaStride * (n_hops - 1) + kMetaF;
        QV_TRY(gather_enqueue(tail.table, n_id, tail.feature_order, std::min(bn[n_hops], tail.capacity_rows), d_n,
                              tail.row_bytes, tail.features, tail.variant, st));
    }
    QV_CUDA(cudaEventSynchronize(s->meta_ready));
    if (n_hops >= 2) {  // more rows above kHeavyDeg than heavy blocks: the rest walked their chain in one warp; grow for the next call
        const int64_t seen = s->h_meta[kMetaStride * (n_hops - 2) + kMetaHeavy];
        static const bool fixed = getenv("QV_HEAVY_FRONT_FIXED") != nullptr;  // A-B switch: keep the initial cap
        if (seen > s->heavy_front && !fixed)
            s->heavy_front = static_cast<int>(std::min<int64_t>(kHeavyListMax, (seen + seen / 4 + 63) / 64 * 64));
    }
    if (s->h_meta[kMetaErr] != 0) {
        *id_error = true;  // the caller redoes the call on the hash path
        s->err_dirty = true;
    }
    return QV_OK;
}

int khop_entry(qv_sampler *s, const int64_t *seeds, int64_t S, const int64_t *sizes, int n_hops, uint64_t rand_seed,
               int64_t *n_id, int64_t *const *edge_buf, int64_t *const *eid_buf, int64_t *out_nodes, int64_t *out_edges,
               qv_stream_t stream, const GatherTail &tail);
}  // namespace

int qv_khop(qv_sampler *s, const int64_t *seeds, int64_t S, const int64_t *sizes, int n_hops, uint64_t rand_seed,
            int64_t *n_id, int64_t *const *edge_buf, int64_t *const *eid_buf, int64_t *out_nodes, int64_t *out_edges,
            qv_stream_t stream)
{
    return khop_entry(s, seeds, S, sizes, n_hops, rand_seed, n_id, edge_buf, eid_buf, out_nodes, out_edges, stream,
                      GatherTail());
}

int qv_khop_gather(qv_sampler *s, const int64_t *seeds, int64_t S, const int64_t *sizes, int n_hops, uint64_t rand_seed,
                   int64_t *n_id, int64_t *const *edge_buf, int64_t *const *eid_buf, const qv_shard_table *table,
                   const int64_t *feature_order, int64_t row_bytes, void *features, int64_t features_rows, int variant,
                   int64_t *out_nodes, int64_t *out_edges, qv_stream_t stream)
{
    QV_REQUIRE(table != nullptr, "qv_khop_gather: table is NULL");
    QV_REQUIRE(features != nullptr || S == 0, "qv_khop_gather: features is NULL");
    QV_REQUIRE(row_bytes > 0, "qv_khop_gather: row_bytes = %lld", (long long)row_bytes);
    GatherTail tail;
    tail.table = table;
    tail.feature_order = feature_order;
    tail.row_bytes = row_bytes;
    tail.features = features;
    tail.capacity_rows = features_rows;
    tail.variant = variant;
    QV_REQUIRE(features_rows >= 0, "qv_khop_gather: features_rows = %lld", (long long)features_rows);
    QV_TRY(khop_entry(s, seeds, S, sizes, n_hops, rand_seed, n_id, edge_buf, eid_buf, out_nodes, out_edges, stream, tail));
    if (out_nodes[n_hops] > features_rows)
        return fail(QV_ERR_UNSUPPORTED, "qv_khop_gather: the frontier has %lld rows, `features` holds %lld: only those were "
                    "gathered (sample results are complete; gather the rest with qv_gather)", (long long)out_nodes[n_hops],
                    (long long)features_rows);
    return QV_OK;
}

namespace
{
int khop_entry(qv_sampler *s, const int64_t *seeds, int64_t S, const int64_t *sizes, int n_hops, uint64_t rand_seed,
               int64_t *n_id, int64_t *const *edge_buf, int64_t *const *eid_buf, int64_t *out_nodes, int64_t *out_edges,
               qv_stream_t stream, const GatherTail &tail)
{
    QV_REQUIRE(s && sizes && out_nodes && out_edges, "qv_khop: NULL argument");
    int64_t bn[QV_MAX_HOPS + 1], be[QV_MAX_HOPS];
    QV_TRY(qv_khop_bounds(S, sizes, n_hops, bn, be));
    for (int h = 0; h <= n_hops; h++) out_nodes[h] = h == 0 ? S : 0;
    for (int h = 0; h < n_hops; h++) out_edges[h] = 0;
    if (S == 0) return QV_OK;
    QV_REQUIRE(seeds && n_id && edge_buf, "qv_khop: NULL array");
    for (int h = 0; h < n_hops; h++) QV_REQUIRE(edge_buf[h] != nullptr || be[h] == 0, "qv_khop: edge_buf[%d] is NULL", h);
    DeviceGuard g(s->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    int64_t max_nodes = 0, max_edges = 0;
    for (int h = 0; h < n_hops; h++) {
        max_nodes = std::max(max_nodes, bn[h]);
        max_edges = std::max(max_edges, be[h]);
    }
    const char *env = getenv("QV_KHOP_REINDEX");  // "hash" forces the per-hop hash table (tests / A-B)
    bool use_map = !(env && env[0] == 'h') && s->n_nodes > 0 && s->n_nodes < (int64_t(1) << 31) &&
                   s->n_nodes <= (int64_t(1) << 30) && bn[n_hops] < (int64_t(1) << 31);
    QV_TRY(ensure_scan(s, bn[n_hops]));
    QV_TRY(s->out_ptr.ensure(static_cast<size_t>(max_nodes) * sizeof(int64_t)));
    int64_t sum_edges = 0;
    for (int h = 0; h < n_hops; h++) sum_edges += be[h];
    QV_TRY(s->nbr.ensure(static_cast<size_t>(std::max<int64_t>(sum_edges, 1)) * sizeof(int64_t)));
    if (use_map && !s->node_map.ptr) {
        if (s->node_map.ensure(static_cast<size_t>(s->n_nodes) * sizeof(MapWord)) != QV_OK) use_map = false;  // no room
        s->map_epoch = 0;
    }
    if (!use_map) QV_TRY(ensure_table(s, bn[n_hops]));
    const char *env_deg = getenv("QV_KHOP_CACHE_DEG");  // "0": re-read indptr per hop instead of caching (A-B switch)
    if (use_map && !(env_deg && env_deg[0] == '0'))
        QV_TRY(s->fr_meta.ensure(static_cast<size_t>(2 * bn[n_hops]) * sizeof(int64_t)));
    else
        s->fr_meta.release();

    bool id_error = false;
    if (use_map && s->fr_meta.ptr && fused_hops_apply(s, S, sizes, n_hops, bn, be)) {
        int64_t max_tiles = 1;
        for (int h = 0; h < n_hops; h++) max_tiles = std::max(max_tiles, (bn[h] + kSampleTile - 1) / kSampleTile);
        const size_t stride = (static_cast<size_t>(kHopCtlFixed + max_tiles) + 7) & ~size_t(7);
        QV_TRY(s->ctl.ensure((kCtlHeader + stride * n_hops) * sizeof(unsigned long long)));
        QV_TRY(s->tgt.ensure(static_cast<size_t>(std::max<int64_t>(max_edges, 1)) * sizeof(int32_t)));
        QV_TRY(s->tile_base.ensure(static_cast<size_t>(max_tiles) * sizeof(int64_t)));
        QV_TRY(khop_run_fused(s, seeds, S, sizes, n_hops, rand_seed, n_id, edge_buf, eid_buf, bn, be, st, &id_error, tail));
    } else
    QV_TRY(khop_run(s, seeds, S, sizes, n_hops, rand_seed, n_id, edge_buf, eid_buf, bn, be, use_map, st, &id_error, tail));
    if (id_error) {
        QV_TRY(ensure_table(s, bn[n_hops]));
        QV_TRY(khop_run(s, seeds, S, sizes, n_hops, rand_seed, n_id, edge_buf, eid_buf, bn, be, false, st, &id_error, tail));
    }
    for (int h = 0; h < n_hops; h++) {
        out_edges[h] = s->h_meta[kMetaStride * h + kMetaE];
        out_nodes[h + 1] = s->h_meta[kMetaStride * h + kMetaF];
    }
    return QV_OK;
}
}  // namespace

int qv_sampler_set_fast(qv_sampler *s, int enabled)
{
    QV_REQUIRE(s != nullptr, "qv_sampler_set_fast: NULL sampler");
    s->fast = enabled != 0;
    return QV_OK;
}

int qv_cal_neighbor_prob(qv_sampler *s, const float *last_prob, float *cur_prob, int64_t n, int k, qv_stream_t stream)
{
    QV_REQUIRE(s && last_prob && cur_prob, "qv_cal_neighbor_prob: NULL argument");
    QV_REQUIRE(n >= 0 && n <= s->n_nodes, "qv_cal_neighbor_prob: n = %lld outside [0, %lld]", (long long)n,
               (long long)s->n_nodes);
    if (n == 0) return QV_OK;
    DeviceGuard g(s->device);
    cal_next_kernel<<<grid_for(n * 32, 256, s->n_sm, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        last_prob, cur_prob, n, k, s->indptr, s->indices);
    QV_CHECK_LAUNCH("cal_next_kernel");
    return QV_OK;
}

}  // extern "C"
