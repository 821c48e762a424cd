// qv_hop_kernels.cuh -- the two kernels a hop of the fused k-hop sampler consists of (included by qv_sample.cu inside
// namespace qv { namespace { ... } }, after the shared helpers: Xorwow, RecipTable, reservoir_hit, cp_async_8, MapWord ...).
//
//   hop_sample_kernel    degree/cap/exclusive-scan + row-wise reservoir sampling in ONE launch
//                        (replaces count_scan_kernel + sample_rows_small_kernel of round 1; reference:
//                         quiver_sample.cu:157-169 + CSRRowWiseSampleKernel cuda_random.cu.hpp:7-69)
//   hop_reindex_kernel   first-occurrence reindex of the hop's sampled ids in ONE co-resident launch with three grid
//                        barriers (replaces map_insert + map_scan + map_emit_all; reference: quiver_sample.cu:202-255,
//                         18-63, 338-351)
// so a 3-hop sample is 6 launches instead of 13, none of them a chained look-back scan over hundreds of tiles.
//
// Why (round-1 profile of the north-star batch, profiles/r2_launches_ns_before.txt): the sampling kernel of a 9 k-row hop
// took 65-117 us because ONE warp walks a 91 k-degree row (2859 dependent generator draws at 30-100 cycles each), the
// 400 k-item last-hop scan took 40 us waiting on its 400-tile look-back chain, and 8 of the 13 launches were 4-20 us
// kernels that do a few hundred ns of work behind a launch + dependency latency.
//
// Heavy rows.  Parity pins lane l of logical warp (b, w) to ONE XORWOW stream serving that warp's rows in order, so the
// draws of a long row are a dependent chain.  What is NOT pinned is who evaluates them: in a "heavy block" (256 threads)
// warp 0 only advances the 32 generators (7 instructions per draw on a ~15-cycle dependency, profiles/micro/xorwow_chain.cu)
// and streams the raw outputs through a kStreamBufs-deep shared-memory ring guarded by named barriers; six tester warps
// (1-3, 5-7) add the Weyl term, evaluate `r mod m < k` (ten independent `%` per batch: the divides pipeline, a reciprocal
// table's cold loads did not) and atomicMax the reservoir; warp 4, which shares warp 0's scheduler, only joins the
// barriers.  Measured: 23.6 cycles per draw in the kernel against 30-100 before (profiles/r2_hop_sample_ns_full.txt).  The
// warps that own such rows are known before the hop starts: every node's degree is recorded when it joins the frontier
// (hop_reindex_kernel appends rows above kHeavyDeg to a per-call list), so heavy blocks sit at the FRONT of the grid and
// start at time zero (longest-first), and the row's regular slot retires when it finds itself listed.
#pragma once

constexpr int kHeavyDeg = 3072;     // rows with more neighbours join the call's heavy list (>= 96 draws per lane)
constexpr int kHeavyListMax = 2048;  // capacity of the call's heavy list
constexpr int kHeavyFrontInit = 256;  // listed rows = heavy blocks at the front of every later hop's sampling grid: a sampler
                                      // starts here and grows towards kHeavyListMax when a call lists more (R-MAT graphs do)
constexpr int kStreamMin = 64;      // draws per lane from which a row of a heavy block is streamed to the tester warps
constexpr int kStreamChunk = 60;    // draws per lane per ring buffer (a multiple of kGenUnroll)
constexpr int kGenUnroll = 10;      // generator steps per loop iteration (a multiple of 5 keeps the state rotation free of moves)
constexpr int kHopTiles = 2;        // 64-row tiles per regular block
constexpr int kHopWarps = kHopTiles * kSampleWarps;  // 8 warps: a heavy block is 1 generator warp + 6 testers + 1 idle
constexpr int kHopThreads = kHopWarps * 32;
constexpr int kTesters = kHopWarps - 2;  // warp 4 shares the generator's scheduler: it only joins the barriers (see stream_test)
constexpr int kStreamBufs = 4;      // ring depth (named barriers 1..4 = full, 5..8 = empty, 9 = all draws evaluated; all of kHopThreads)

// ---- control words of one fused k-hop call (zeroed by one memset) ----------------------------------------------------
//   ctl[0]                 number of heavy rows listed so far (may exceed the call's cap n_front: entries beyond it are dropped)
//   ctl[1 .. 1+max)        row index (= local id, stable across hops) of each listed row
//   per hop h, at ctl + kCtlHeader + h * stride:
//     [0] tile ticket of the sampling kernel   [1] grid-barrier counter of the reindex kernel
//     [2 .. 2+2*kReindexMaxBlocks)  per-block first-occurrence counts, then per-block next-hop entry counts
//     [2+2*kReindexMaxBlocks ..)    hop 0 only: sampling tile descriptors
constexpr int kCtlHeader = 1 + kHeavyListMax + 7;  // 2056 words: keeps the per-hop regions 64-byte aligned
constexpr int kReindexMaxBlocks = 512;
constexpr int kHopCtlFixed = 2 + 2 * kReindexMaxBlocks;

__device__ int g_hop_debug = 0;  // QV_HOP_DEBUG=1: device-side phase timing printed from the kernels (diagnostics only)
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// bar.sync / bar.arrive are the .aligned forms: every thread of the warp must execute them together, and independent
// thread scheduling does not promise that a warp has reconverged after the divergent reservoir updates in front of them
// (compute-sanitizer synccheck flagged exactly that) -- hence the __syncwarp() first.
__device__ __forceinline__ void named_bar_sync(int id, int count)
{
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count)
{
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

struct HopSampleArgs {
    const int64_t *indptr, *indices;
    int64_t n_nodes;
    const int64_t *seeds;  // the hop's rows: seeds (hop 0) or the frontier so far (n_id)
    int64_t S_arg;
    const int64_t *d_S;
    int k;
    const uint32_t *rng_states;
    RecipTable rt;
    int64_t *out;        // sampled neighbour ids [E]
    int32_t *tgt;        // row (target) index of every sampled edge [E]
    int64_t *eid_out;    // optional edge ids [E]
    const int64_t *edge_ids;
    const int64_t *cached_start, *cached_deg;  // CSR row of every frontier node (null on hop 0)
    const int64_t *tile_base;    // hops >= 1: every tile's output offset, computed by the previous hop's reindex kernel
    MapWord *node_map;           // non-null: the sampled ids (hop 0: the seeds too) enter the node map right here
    unsigned int epoch_hi;
    int64_t item_base;           // item index of the hop's first output (hop 0: number of seeds)
    int64_t *d_err;
    unsigned long long *desc;    // hop 0 only: tile descriptors (decoupled look-back): flag << 62 | value
    unsigned long long *ticket;  // null: tile = block index (the whole grid is resident at once)
    unsigned long long *heavy;   // ctl[0..]: count + listed rows
    int64_t *d_E;                // out: number of sampled edges of the hop
    int n_front;                 // heavy blocks at the front of the grid = the call's list cap (0 on hop 0: the list is still empty)
    int release_early;           // the grid leaves room on every SM: the reindex kernel's blocks may become resident at once
};

struct __align__(16) TileSmem {
    int64_t start[kSampleTile];
    uint32_t deg[kSampleTile];
    uint32_t excl[kSampleTile];  // tile-local exclusive offset of each row's entries
    uint32_t total;
    int tile;  // broadcast slot for the ticket / the listed row
};

struct __align__(16) StreamSmem {
    uint32_t buf[kStreamBufs][kStreamChunk][32];  // raw generator words v4 (the Weyl term is added by the testers)
    uint32_t dchunk[kStreamBufs][32];             // each lane's Weyl counter at the start of the chunk
};

// Loads the CSR rows of tile b (by the 64 threads starting at `first_thread`), caps the counts and scans them (by the first
// warp of those).  Called by ALL threads of the block: two block barriers.
__device__ __forceinline__ void tile_prologue(const HopSampleArgs &a, TileSmem &sm, int64_t S, int64_t b, bool publish,
                                              int first_thread)
{
    const int t = static_cast<int>(threadIdx.x) - first_thread;
    if (t >= 0 && t < kSampleTile) {
        const int64_t r = b * kSampleTile + t;
        int64_t start = 0, deg = 0;
        if (r < S) {
            if (a.cached_deg) {
                start = a.cached_start[r];
                deg = a.cached_deg[r];
            } else {
                const int64_t node = a.seeds[r];
                if (node >= 0 && node < a.n_nodes) {
                    start = a.indptr[node];
                    deg = a.indptr[node + 1] - start;
                }
            }
        }
        sm.start[t] = start;
        sm.deg[t] = static_cast<uint32_t>(min(deg, static_cast<int64_t>(0xffffffffu)));
        if (publish && a.node_map && !a.cached_deg && r < S) {  // hop 0: the seeds enter the node map
            const int64_t node = a.seeds[r];
            if (node >= 0 && node < a.n_nodes)
                atomicMin(&a.node_map[node], map_word(a.epoch_hi, kMapCand + static_cast<unsigned int>(r)));
            else
                *a.d_err = 1;
        }
    }
    __syncthreads();
    if (t >= 0 && t < 32) {  // exclusive scan of the 64 capped counts, two per lane
        const uint32_t kk = static_cast<uint32_t>(a.k);
        const uint32_t c0 = min(sm.deg[2 * t], kk), c1 = min(sm.deg[2 * t + 1], kk);
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, off);
            if (t >= off) incl += u;
        }
        sm.excl[2 * t] = incl - c0 - c1;
        sm.excl[2 * t + 1] = incl - c1;
        if (t == 31) sm.total = incl;
    }
    __syncthreads();
}

// Exclusive prefix of tile b over the tiles in front of it (decoupled look-back, executed by one warp).
__device__ __forceinline__ long long tile_lookback(const unsigned long long *desc, int64_t b, int lane)
{
    long long run = 0;
    int64_t look = b - 1;
    while (look >= 0) {
        const int64_t idx = look - lane;
        unsigned long long w = idx >= 0 ? ld_volatile_u64(desc + idx) : kFlagPrefix;
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
    return run;
}

// One XORWOW step without the Weyl term (xorwow_next = this + `d += 362437; return v4 + d`).
__device__ __forceinline__ uint32_t xorwow_raw(Xorwow &s)
{
    const uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1;
    s.v1 = s.v2;
    s.v2 = s.v3;
    s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    return s.v4;
}

// Generator side of a streamed row: `rem` = this lane's draws in the row, n_chunks = chunks of lane 0 (the most).
__device__ __noinline__ Xorwow stream_generate(Xorwow rng, uint32_t rem, uint32_t n_chunks, StreamSmem *ss, int lane)
{
    const bool prof = g_hop_debug == 4;
    long long t_wait = 0, t_all = prof ? clock64() : 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        const int p = c % kStreamBufs;
        const long long w0 = prof ? clock64() : 0;
        named_bar_sync(1 + kStreamBufs + p, kHopThreads);  // the testers are done with this buffer (or primed it)
        if (prof) t_wait += clock64() - w0;
        ss->dchunk[p][lane] = rng.d;
        const uint32_t t0 = c * kStreamChunk;
        if (t0 + kStreamChunk <= rem) {
            // rolled in steps of kGenUnroll: the whole loop stays in the instruction cache next to the tester sharing this
            // warp's scheduler (a fully unrolled chunk is 7 KB of straight-line code)
            uint32_t *dst = &ss->buf[p][0][lane];
#pragma unroll 1
            for (int t = 0; t < kStreamChunk; t += kGenUnroll) {
#pragma unroll
                for (int u = 0; u < kGenUnroll; u++) dst[(t + u) * 32] = xorwow_raw(rng);
            }
            rng.d += kStreamChunk * 362437u;
        } else {
            const uint32_t left = rem > t0 ? rem - t0 : 0;
            for (uint32_t t = 0; t < left; t++) ss->buf[p][t][lane] = xorwow_raw(rng);
            rng.d += left * 362437u;
        }
        named_bar_arrive(1 + p, kHopThreads);
    }
    if (prof && lane == 0 && n_chunks > 20)
        printf("[stream generator] %u chunks of %d rounds: %lld cycles, %lld of them waiting for a free buffer\n", n_chunks,
               kStreamChunk, (long long)(clock64() - t_all), t_wait);
    return rng;
}

// Tester side (warps 1..7 of a heavy block, q = 0..6): the same sequence of streamed rows, derived from the tile's degrees.
__device__ __noinline__ void stream_test(const HopSampleArgs a, const TileSmem *sm, int w, int q, int lane,
                                         uint32_t *slots_w, uint32_t kcap, StreamSmem *ss)
{
    const uint32_t kk = static_cast<uint32_t>(a.k);
    const uint32_t first = kk + lane;
    for (int i = 0; i < kRowsPerWarp; i++) {
        const uint32_t d = sm->deg[w + kSampleWarps * i];
        if (d <= kk) continue;
        const uint32_t rem0 = (d - kk + 31) >> 5;
        if (rem0 < kStreamMin) continue;
        uint32_t *srow = slots_w + static_cast<size_t>(i) * kcap;
        const uint32_t n_chunks = (rem0 + kStreamChunk - 1) / kStreamChunk;
        // This warp evaluates rounds q, q + 7, ... of every chunk with the plain `%` operator: no table walk (a streamed row
        // would read its own cold stretch of the reciprocal table, and holding two chunks of 64-bit reciprocals in registers
        // spills under the kernel's 64-register budget -- both measured slower).  All shared-memory reads of the chunk come
        // BEFORE the first reservoir update: the updates are shared-memory atomics through a generic pointer and the compiler
        // keeps every later load behind a possible earlier store, which serialises the eleven divisions otherwise.
        constexpr int kPer = (kStreamChunk + kTesters - 1) / kTesters;
#pragma unroll
        for (int p = 0; p < kStreamBufs; p++)  // every buffer starts empty
            if (static_cast<uint32_t>(p) < n_chunks) named_bar_arrive(1 + kStreamBufs + p, kHopThreads);
        for (uint32_t c = 0; c < n_chunks; c++) {
            const int p = c % kStreamBufs;
            named_bar_sync(1 + p, kHopThreads);
            const uint32_t d0 = ss->dchunk[p][lane];
            uint32_t rr[kPer], num[kPer];
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                const int t = q + kTesters * u;
                rr[u] = (q >= 0 && t < kStreamChunk) ? ss->buf[p][t][lane] + d0 + static_cast<uint32_t>(t + 1) * 362437u : 0u;
            }
            if (c + kStreamBufs < n_chunks) named_bar_arrive(1 + kStreamBufs + p, kHopThreads);  // the buffer is free again
            if (q < 0) continue;  // the warp on the generator's scheduler stays out of its way
            unsigned int hit = 0;
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                const uint32_t idx = first + 32u * (c * kStreamChunk + q + kTesters * u);
                num[u] = rr[u] % (idx + 1);
                if (q + kTesters * u < kStreamChunk && idx < d && num[u] < kk) hit |= 1u << u;
            }
            if (hit) {
#pragma unroll
                for (int u = 0; u < kPer; u++)
                    if (hit >> u & 1u) atomicMax(&srow[num[u]], first + 32u * (c * kStreamChunk + q + kTesters * u));
            }
        }
    }
    named_bar_sync(1 + 2 * kStreamBufs, kHopThreads);  // every streamed draw has been evaluated: the generator may read the reservoirs
}

// The rows of logical warp (b, w), executed by one physical warp whose shared-memory slice is (stage_w, slots_w, rowof_w,
// pre_w).  kStream: this is the generator warp of a heavy block (rows with >= kStreamMin draws per lane are streamed).
template <bool kStream, int kHub>
__device__ __forceinline__ void warp_rows(const HopSampleArgs &a, const TileSmem &sm, int64_t S, int64_t n_tiles, int64_t b,
                                          int w, int lane, int64_t *stage_w, uint32_t *slots_w, uint8_t *rowof_w,
                                          uint16_t *pre_w, StreamSmem *ss)
{
    const uint32_t kk = static_cast<uint32_t>(a.k);
    const uint32_t kcap = a.k > 0 ? kk : 1u;
    const uint32_t per_warp = kRowsPerWarp * kcap;
    Xorwow rng;
    {
        const uint32_t *p = a.rng_states + static_cast<size_t>(b) * kRngStateWords * kRngBlockThreads + (w * 32 + lane);
        rng.d = p[0 * kRngBlockThreads];
        rng.v0 = p[1 * kRngBlockThreads];
        rng.v1 = p[2 * kRngBlockThreads];
        rng.v2 = p[3 * kRngBlockThreads];
        rng.v3 = p[4 * kRngBlockThreads];
        rng.v4 = p[5 * kRngBlockThreads];
    }
    // lane i < 16 lays out the entries of row i (tile row w + 4 i) back to back in the warp's list
    uint32_t n_entries;
    {
        const uint32_t cnt = lane < kRowsPerWarp ? min(sm.deg[w + kSampleWarps * lane], kk) : 0u;
        uint32_t incl = cnt;
#pragma unroll
        for (int off = 1; off < kRowsPerWarp; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        n_entries = __shfl_sync(0xffffffffu, incl, kRowsPerWarp - 1);
        if (lane < kRowsPerWarp) {
            pre_w[lane] = static_cast<uint16_t>(incl - cnt);
            for (uint32_t j = 0; j < cnt; j++) rowof_w[incl - cnt + j] = static_cast<uint8_t>(lane);
        }
    }
    {  // every reservoir starts as 0..k-1 (e % kcap by multiply-high: exact for e < 2^16)
        const uint32_t inv_k = kcap > 1 ? 0xFFFFFFFFu / kcap + 1u : 0u;
        for (uint32_t e = lane; e < per_warp; e += 32) slots_w[e] = kcap > 1 ? e - __umulhi(e, inv_k) * kcap : 0u;
    }
    __syncwarp();

    // verbatim rows: their ids can start travelling now
    for (uint32_t e = lane; e < n_entries; e += 32) {
        const int i = rowof_w[e];
        const int tr = w + kSampleWarps * i;
        if (sm.deg[tr] <= kk) cp_async_8(&stage_w[e], a.indices + sm.start[tr] + (e - pre_w[i]));
    }

    // this lane's generator stream, row after row, no synchronisation
    {
        const uint32_t first = kk + lane;
        const unsigned long long *tab = a.rt.recip + 1;  // tab[idx] = recip[idx + 1]
        const uint32_t tab_n = a.rt.n > 0 ? a.rt.n - 1 : 0;
        for (int i = 0; i < kRowsPerWarp; i++) {
            const uint32_t d = sm.deg[w + kSampleWarps * i];
            uint32_t *srow = slots_w + static_cast<size_t>(i) * kcap;
            if (d > kk && d <= tab_n && (d - kk + 31) >> 5 < kHub) {
                // the common row: <= kHub-1 draws per lane, fully unrolled and predicated (warp-uniform test above)
#pragma unroll
                for (int t = 0; t < kHub - 1; t++) {
                    const uint32_t idx = first + 32u * t;
                    if (idx < d) reservoir_hit(tab[idx], xorwow_next(rng), idx + 1, kk, idx, srow);
                }
                continue;
            }
            if (d <= kk) continue;
            if (kStream && (d - kk + 31) >> 5 >= kStreamMin) {  // warp-uniform: the testers take the same decision
                const uint32_t rem_l = d > first ? (d - first + 31) >> 5 : 0;
                const uint32_t rem0 = (d - kk + 31) >> 5;
                rng = stream_generate(rng, rem_l, (rem0 + kStreamChunk - 1) / kStreamChunk, ss, lane);
                continue;
            }
            if (d <= first) continue;
            uint32_t rem = (d - first + 31) >> 5, idx = first;
            if (rem >= kHub && idx + 64 * kHub < tab_n) {
                unsigned long long M[kHub];
#pragma unroll
                for (int u = 0; u < kHub; u++) M[u] = tab[idx + 32 * u];
                while (true) {
                    unsigned long long N[kHub];
                    const bool more = rem >= 2 * kHub && idx + 96 * kHub < tab_n;
                    prefetch_l1(tab + min(idx + 32u * kHub * 6u, tab_n - 1));
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
        }
    }
    if (kStream) named_bar_sync(1 + 2 * kStreamBufs, kHopThreads);  // the tester warps have evaluated every streamed draw
    __syncwarp();

    // sampled rows: fetch the chosen positions; then one wait and one coalesced write-out of the whole list
    for (uint32_t e = lane; e < n_entries; e += 32) {
        const int i = rowof_w[e];
        const int tr = w + kSampleWarps * i;
        if (sm.deg[tr] > kk) cp_async_8(&stage_w[e], a.indices + sm.start[tr] + slots_w[static_cast<size_t>(i) * kcap + (e - pre_w[i])]);
    }
    // the tile's output offset: hops >= 1 read it (the previous hop's reindex kernel scanned the capped degrees of the
    // whole frontier); hop 0 sums the tiles in front of it (they published their totals when they started)
    long long base;
    if (a.tile_base) {
        base = a.tile_base[b];
    } else {
        base = tile_lookback(a.desc, b, lane);
        if (lane == 0) {
            st_volatile_u64(a.desc + b, kFlagPrefix | (static_cast<unsigned long long>(base + sm.total) & kValueMask));
            if (b == n_tiles - 1) *a.d_E = base + sm.total;
        }
    }
    cp_async_wait_all();
    for (uint32_t e = lane; e < n_entries; e += 32) {
        const int i = rowof_w[e];
        const int tr = w + kSampleWarps * i;
        const uint32_t j = e - pre_w[i];
        const int64_t dst = base + sm.excl[tr] + j;
        const int64_t id = stage_w[e];
        a.out[dst] = id;
        a.tgt[dst] = static_cast<int32_t>(b * kSampleTile + tr);
        if (a.node_map) {
            if (static_cast<uint64_t>(id) < static_cast<uint64_t>(a.n_nodes))
                atomicMin(&a.node_map[id], map_word(a.epoch_hi, kMapCand + static_cast<unsigned int>(a.item_base + dst)));
            else
                *a.d_err = 1;
        }
        if (a.eid_out) {
            const uint32_t pos = sm.deg[tr] > kk ? slots_w[static_cast<size_t>(i) * kcap + j] : j;
            const int64_t p = sm.start[tr] + pos;
            a.eid_out[dst] = a.edge_ids ? a.edge_ids[p] : p;
        }
    }
}

template <int kHub, int kMinBlocks>
__global__ void __launch_bounds__(kHopThreads, kMinBlocks) hop_sample_kernel(const __grid_constant__ HopSampleArgs a)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ TileSmem sm[kHopTiles];
    __shared__ StreamSmem ss;
    __shared__ uint16_t pre_sh[kHopWarps][kRowsPerWarp];
    pdl_wait();  // everything this hop reads (frontier, CSR rows, sizes, heavy list) is the previous kernel's output
    if (a.release_early) pdl_release();
    const int64_t S = dev_size(a.S_arg, a.d_S);
    const int64_t n_tiles = (S + kSampleTile - 1) / kSampleTile;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const uint32_t kcap = a.k > 0 ? static_cast<uint32_t>(a.k) : 1u;
    const uint32_t per_warp = kRowsPerWarp * kcap;
    int64_t *stage_w = reinterpret_cast<int64_t *>(dyn_smem) + static_cast<size_t>(wp) * per_warp;
    uint32_t *slots_w = reinterpret_cast<uint32_t *>(dyn_smem + static_cast<size_t>(kHopWarps) * per_warp * 8) +
                        static_cast<size_t>(wp) * per_warp;
    uint8_t *rowof_w = dyn_smem + static_cast<size_t>(kHopWarps) * per_warp * 12 + static_cast<size_t>(wp) * per_warp;

    if (static_cast<int>(blockIdx.x) < a.n_front) {
        // ---- heavy block: one listed row's logical warp; warp 0 generates, warps 1..7 test ------------------------------
        const unsigned long long n_listed = min(a.heavy[0], static_cast<unsigned long long>(a.n_front));
        if (blockIdx.x >= n_listed) return;
        const int64_t r = static_cast<int64_t>(a.heavy[1 + blockIdx.x]);
        const int64_t b = r >> 6;
        const int w = static_cast<int>(r & 3);
        bool dup = false;  // an earlier entry of the same warp serves it
        for (unsigned int j = threadIdx.x; j < blockIdx.x; j += blockDim.x) {
            const int64_t o = static_cast<int64_t>(a.heavy[1 + j]);
            dup |= (o >> 6) == b && (o & 3) == w;
        }
        if (__syncthreads_or(dup) || r >= S) return;
        const unsigned long long t0 = g_hop_debug ? global_ns() : 0;
        tile_prologue(a, sm[0], S, b, false, 0);
        if (wp == 0) {
            warp_rows<true, kHub>(a, sm[0], S, n_tiles, b, w, lane, stage_w, slots_w, rowof_w, pre_sh[0], &ss);
            if (g_hop_debug && lane == 0) {
                unsigned int rounds = 0, big = 0;
                for (int i = 0; i < kRowsPerWarp; i++) {
                    const uint32_t d = sm[0].deg[w + kSampleWarps * i];
                    if (d > static_cast<uint32_t>(a.k)) rounds += (d - a.k + 31) >> 5;
                    big = max(big, d);
                }
                printf("[hop_sample k=%d] heavy block %d: %u rounds (max deg %u) in %.1f us\n", a.k, blockIdx.x, rounds, big,
                       (global_ns() - t0) * 1e-3);
            }
        } else
            stream_test(a, &sm[0], w, wp < kSampleWarps ? wp - 1 : (wp == kSampleWarps ? -1 : wp - 2), lane,
                        slots_w - static_cast<size_t>(wp) * per_warp, kcap, &ss);
        return;
    }

    // ---- regular block: kHopTiles consecutive tiles of 64 rows, four warps each ---------------------------------------------
    int64_t blk = static_cast<int64_t>(blockIdx.x) - a.n_front;
    if (a.ticket) {  // grids larger than the device holds at once: tiles in dispatch order, so look-back never waits on a
        if (threadIdx.x == 0) sm[0].tile = static_cast<int>(atomicAdd(a.ticket, 1ull));  // tile whose block has not started
        __syncthreads();
        blk = sm[0].tile;
        __syncthreads();
    }
    if (blk * kHopTiles >= n_tiles) return;
    const int ts = wp / kSampleWarps;  // this warp's tile slot
    const int64_t b = blk * kHopTiles + ts;
    // both tiles' CSR rows are loaded at once (threads 0..63 / 128..191), scanned by warps 0 / 4
    {
        const int t = static_cast<int>(threadIdx.x) - ts * (kSampleWarps * 32);
        if (t < kSampleTile) {
            const int64_t r = b * kSampleTile + t;
            int64_t start = 0, deg = 0;
            if (r < S) {
                if (a.cached_deg) {
                    start = a.cached_start[r];
                    deg = a.cached_deg[r];
                } else {
                    const int64_t node = a.seeds[r];
                    if (node >= 0 && node < a.n_nodes) {
                        start = a.indptr[node];
                        deg = a.indptr[node + 1] - start;
                        if (a.node_map)  // hop 0: the seeds enter the node map
                            atomicMin(&a.node_map[node], map_word(a.epoch_hi, kMapCand + static_cast<unsigned int>(r)));
                    } else if (a.node_map) {
                        *a.d_err = 1;
                    }
                }
            }
            sm[ts].start[t] = start;
            sm[ts].deg[t] = static_cast<uint32_t>(min(deg, static_cast<int64_t>(0xffffffffu)));
        }
        __syncthreads();
        if (t < 32) {
            const uint32_t kk = static_cast<uint32_t>(a.k);
            const uint32_t c0 = min(sm[ts].deg[2 * t], kk), c1 = min(sm[ts].deg[2 * t + 1], kk);
            uint32_t incl = c0 + c1;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, off);
                if (t >= off) incl += u;
            }
            sm[ts].excl[2 * t] = incl - c0 - c1;
            sm[ts].excl[2 * t + 1] = incl - c1;
            if (t == 31) {
                sm[ts].total = incl;
                if (!a.tile_base && b < n_tiles)  // hop 0: the tile's count is public from now on (look-back at write-out)
                    st_volatile_u64(a.desc + b, (b == 0 ? kFlagPrefix : kFlagAgg) | (static_cast<unsigned long long>(incl) & kValueMask));
            }
        }
        __syncthreads();
    }
    if (b >= n_tiles) return;
    const int w = wp % kSampleWarps;
    if (a.n_front > 0) {
        // does a heavy block serve this warp?  (only warps owning a listed-size row need to look)
        const bool big = lane < kRowsPerWarp && sm[ts].deg[w + kSampleWarps * lane] > kHeavyDeg;
        if (__any_sync(0xffffffffu, big)) {
            const unsigned int n_listed = static_cast<unsigned int>(min(a.heavy[0], static_cast<unsigned long long>(a.n_front)));
            bool found = false;
            for (unsigned int j = lane; j < n_listed; j += 32) {
                const int64_t o = static_cast<int64_t>(a.heavy[1 + j]);
                found |= (o >> 6) == b && (o & 3) == w;
            }
            if (__any_sync(0xffffffffu, found)) return;
        }
    }
    warp_rows<false, kHub>(a, sm[ts], S, n_tiles, b, w, lane, stage_w, slots_w, rowof_w, pre_sh[wp], nullptr);
}

// ------------------------------------------------------------------------------------------------------------------
// hop_reindex_kernel: items = [prefix (hop 0: the seeds) | the hop's sampled ids], one tile of 512 * kItems items per block,
// every block resident (the host bounds the grid by the occupancy), grid barriers between the phases:
//   1  atomicMin the item index into the epoch-tagged node map (see MapWord) -- skipped when the sampling kernel already
//      did it at write-out time (insert_done)                                           -- barrier --
//   2  an item is a first occurrence iff the map still holds ITS index; count per block  -- barrier --
//      block offset = sum of the earlier blocks' counts (<= 512 values, one load each); new nodes get consecutive local
//      ids in item order, enter the frontier, have their CSR row recorded for the next hop, join the heavy list if big
//                                                                                        -- barrier --
//   3  col[e] = local id of item e (map look-up, L2-hot), row[e] = its target; written as the PyG edge_index [2, E]
//   4  (not on the last hop) the NEXT hop's output offsets: sum of min(degree, k_next) per 64-row tile of the new frontier,
//      scanned over the blocks                                                           -- barrier --
//      so the next sampling kernel needs no count kernel, no look-back chain and knows its edge count up front.
// The ids stay in registers from phase 1 to phase 3.
// ------------------------------------------------------------------------------------------------------------------
struct HopReindexArgs {
    const int64_t *prefix;
    int64_t P_arg;
    const int64_t *nbr;
    const int64_t *d_E;
    MapWord *map;
    unsigned int epoch_hi;
    int insert_done;
    int64_t n_nodes;
    const int64_t *d_F_prev;  // null on hop 0 (the frontier starts empty)
    int64_t *frontier;
    int64_t *d_F;
    int64_t *d_next_S;
    const int64_t *indptr;
    int64_t *fr_start, *fr_deg;  // null on the last hop
    unsigned long long *heavy;
    unsigned int heavy_cap;   // entries the list takes in this call (= the sampling kernels' n_front)
    int64_t *d_heavy_seen;    // rows above kHeavyDeg seen so far in this call (listed or not): the host sizes the next call's cap
    const int32_t *tgt;
    int64_t *edge_buf;  // [col (E) | row (E)]
    unsigned long long *bar;
    unsigned long long *agg;   // [2][kReindexMaxBlocks]: first-occurrence counts, then next-hop entry counts
    int64_t *tile_base_next;   // out (not on the last hop): output offset of every 64-row tile of the next hop
    int64_t *d_E_next;         // out: the next hop's edge count
    int64_t k_next;
    int64_t *d_err;
};

constexpr int kReindexThreads = 512;
constexpr int kReindexTilesPerBlock = 256;  // next-hop tiles one block can scan (bounds the fused path: see plan_reindex)

__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1ull);
        while (ld_volatile_u64(bar) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

template <int kItems>
__global__ void __launch_bounds__(kReindexThreads, 2) hop_reindex_kernel(const __grid_constant__ HopReindexArgs a)
{
    constexpr int kWarps = kReindexThreads / 32;
    __shared__ uint32_t wt[kItems][kWarps];    // first occurrences per (item row, warp)
    __shared__ uint32_t wex[kItems][kWarps];   // exclusive over the warps of a row
    __shared__ uint32_t rowtot[kItems];
    __shared__ long long red[2][kWarps];
    __shared__ long long block_base_sh, grand_total_sh;
    __shared__ uint32_t tsum[kReindexTilesPerBlock];
    pdl_wait();
    pdl_release();  // the next kernel's blocks may become resident (they park in their own wait): this grid already is
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t P = a.prefix ? a.P_arg : 0, E = *a.d_E;
    const int64_t n = P + E;
    const long long F_prev = a.d_F_prev ? *a.d_F_prev : 0;
    const unsigned long long G = gridDim.x;
    unsigned long long bar_target = 0;
    // The grid was sized for the hop's static BOUND; the items that exist are dealt evenly over all of its blocks (in order,
    // whole 512-item rows), so every SM has requests in flight -- a block holding 8 random reads per thread is limited by its
    // SM's outstanding-miss capacity (measured: 11 us for the 380 k map reads of a last hop when 93 of 220 blocks held them).
    const int64_t rows_per_block = ((n + static_cast<int64_t>(G) - 1) / static_cast<int64_t>(G) + kReindexThreads - 1) / kReindexThreads;
    const int64_t tile_base = static_cast<int64_t>(blockIdx.x) * rows_per_block * kReindexThreads;
    const int64_t tile_end = min(n, tile_base + rows_per_block * kReindexThreads);  // rows_per_block <= kItems (host bound)

    // ---- phase 1: the ids (all loads first: they are independent), then insert -----------------------------------------------
    long long key[kItems];
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        const int64_t i = tile_base + j * kReindexThreads + threadIdx.x;
        key[j] = -1;
        if (i < tile_end) key[j] = i < P ? a.prefix[i] : a.nbr[i - P];
    }
    bool bad = false;
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        const int64_t i = tile_base + j * kReindexThreads + threadIdx.x;
        if (i < tile_end) {
            if (static_cast<uint64_t>(key[j]) < static_cast<uint64_t>(a.n_nodes)) {
                if (!a.insert_done) atomicMin(&a.map[key[j]], map_word(a.epoch_hi, kMapCand + static_cast<unsigned int>(i)));
            } else {
                key[j] = -1;
                bad = true;
            }
        }
    }
    if (bad) *a.d_err = 1;
    if (!a.insert_done) grid_barrier(a.bar, bar_target += G);

    // ---- phase 2: first occurrences, counted in item order --------------------------------------------------------------------
    // col[j]: the item's local id if its node is already in the frontier (payload < 2^31), else the candidate word -- such an
    // item is resolved after the assign phase (it is a first occurrence, or a repeat of one inside this hop)
    unsigned int col[kItems];
    {
        MapWord word[kItems];
#pragma unroll
        for (int j = 0; j < kItems; j++) word[j] = key[j] >= 0 ? __ldcg(&a.map[key[j]]) : 0ull;
#pragma unroll
        for (int j = 0; j < kItems; j++) col[j] = static_cast<unsigned int>(word[j]);
    }
    unsigned int first_mask = 0, lt_count[kItems];
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        const int64_t i = tile_base + j * kReindexThreads + threadIdx.x;
        const bool first = key[j] >= 0 && col[j] == kMapCand + static_cast<unsigned int>(i);
        const unsigned int bal = __ballot_sync(0xffffffffu, first);
        if (first) first_mask |= 1u << j;
        lt_count[j] = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) wt[j][warp] = __popc(bal);
    }
    __syncthreads();
    if (warp < kItems) {  // warp j scans row j's 16 warp counts
        const uint32_t v = lane < kWarps ? wt[warp][lane] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int off = 1; off < kWarps; off <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += u;
        }
        if (lane < kWarps) wex[warp][lane] = incl - v;
        if (lane == kWarps - 1) rowtot[warp] = incl;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
#pragma unroll
        for (int j = 0; j < kItems; j++) tot += rowtot[j];
        __stcg(a.agg + blockIdx.x, tot);
    }
    grid_barrier(a.bar, bar_target += G);

    {  // offset of this block = sum of the counts of the blocks in front of it; everybody also learns the grand total
        long long s = 0, all = 0;
        for (unsigned int x = threadIdx.x; x < gridDim.x; x += kReindexThreads) {
            const long long v = static_cast<long long>(__ldcg(a.agg + x));
            all += v;
            if (x < blockIdx.x) s += v;
        }
        s = warp_sum_i64(s);
        all = warp_sum_i64(all);
        if (lane == 0) {
            red[0][warp] = s;
            red[1][warp] = all;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long t = 0, u = 0;
#pragma unroll
            for (int x = 0; x < kWarps; x++) {
                t += red[0][x];
                u += red[1][x];
            }
            block_base_sh = t;
            grand_total_sh = u;
            if (blockIdx.x == gridDim.x - 1) {
                *a.d_F = F_prev + u;
                *a.d_next_S = F_prev + u;
            }
        }
        __syncthreads();
    }
    const long long F = F_prev + grand_total_sh;
    {
        long long row_base = F_prev + block_base_sh;
        constexpr int kBatch = kItems < 4 ? kItems : 4;  // CSR rows of the new nodes: four items' loads in flight at a time
#pragma unroll
        for (int j0 = 0; j0 < kItems; j0 += kBatch) {
            long long rs[kBatch], rd[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; u++) {
                rs[u] = rd[u] = 0;
                if (a.fr_start && (first_mask >> (j0 + u) & 1u)) {
                    rs[u] = a.indptr[key[j0 + u]];
                    rd[u] = a.indptr[key[j0 + u] + 1];
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; u++) {
                const int j = j0 + u;
                if (first_mask >> j & 1u) {
                    const long long local = row_base + wex[j][warp] + lt_count[j];
                    col[j] = static_cast<unsigned int>(local);
                    a.frontier[local] = key[j];
                    __stcg(&a.map[key[j]], map_word(a.epoch_hi, static_cast<unsigned int>(local)));
                    if (a.fr_start) {
                        const long long deg = rd[u] - rs[u];
                        a.fr_start[local] = rs[u];
                        a.fr_deg[local] = deg;
                        if (deg > kHeavyDeg) {
                            const unsigned long long at = atomicAdd(a.heavy, 1ull);
                            if (at < a.heavy_cap) a.heavy[1 + at] = static_cast<unsigned long long>(local);
                        }
                    }
                }
                row_base += rowtot[j];
            }
        }
    }
    grid_barrier(a.bar, bar_target += G);

    // ---- phase 3: edge_index ---------------------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < kItems; j++)  // only repeats of a node first seen in THIS hop still hold a candidate word: look again
        if (key[j] >= 0 && (col[j] & kMapCand)) col[j] = static_cast<unsigned int>(__ldcg(&a.map[key[j]]));
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        const int64_t i = tile_base + j * kReindexThreads + threadIdx.x;
        if (i >= P && i < tile_end) {
            const int64_t e = i - P;
            a.edge_buf[e] = key[j] >= 0 ? static_cast<int64_t>(col[j] & 0x7FFFFFFFu) : 0;
            a.edge_buf[E + e] = a.tgt[e];
        }
    }

    // ---- phase 4: the next hop's tile offsets -------------------------------------------------------------------------------
    if (a.tile_base_next) {
        const int64_t n_tiles = (F + kSampleTile - 1) / kSampleTile;
        const int64_t tpb = (n_tiles + gridDim.x - 1) / gridDim.x;  // <= kReindexTilesPerBlock (host-checked bound)
        const int64_t t_lo = min(n_tiles, static_cast<int64_t>(blockIdx.x) * tpb), t_hi = min(n_tiles, t_lo + tpb);
        for (int64_t t = t_lo + warp; t < t_hi; t += kWarps) {  // a warp sums one tile: 64 rows, two per lane
            const int64_t r0 = t * kSampleTile + 2 * lane;
            long long c = 0;
            if (r0 < F) c += min(static_cast<long long>(__ldcg(a.fr_deg + r0)), static_cast<long long>(a.k_next));
            if (r0 + 1 < F) c += min(static_cast<long long>(__ldcg(a.fr_deg + r0 + 1)), static_cast<long long>(a.k_next));
            c = warp_sum_i64(c);
            if (lane == 0) tsum[t - t_lo] = static_cast<uint32_t>(c);
        }
        __syncthreads();
        long long mine = 0;
        for (int64_t t = threadIdx.x; t < t_hi - t_lo; t += kReindexThreads) mine += tsum[t];
        mine = warp_sum_i64(mine);
        if (lane == 0) red[0][warp] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long t = 0;
#pragma unroll
            for (int x = 0; x < kWarps; x++) t += red[0][x];
            __stcg(a.agg + kReindexMaxBlocks + blockIdx.x, static_cast<unsigned long long>(t));
        }
        grid_barrier(a.bar, bar_target += G);
        long long s = 0, all = 0;
        for (unsigned int x = threadIdx.x; x < gridDim.x; x += kReindexThreads) {
            const long long v = static_cast<long long>(__ldcg(a.agg + kReindexMaxBlocks + x));
            all += v;
            if (x < blockIdx.x) s += v;
        }
        s = warp_sum_i64(s);
        all = warp_sum_i64(all);
        if (lane == 0) {
            red[0][warp] = s;
            red[1][warp] = all;
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of this block's tile sums (<= 256: eight per lane), offset by the earlier blocks
            long long base = 0, tot = 0;
#pragma unroll
            for (int x = 0; x < kWarps; x++) {
                base += red[0][x];
                tot += red[1][x];
            }
            if (lane == 0 && blockIdx.x == gridDim.x - 1) {
                *a.d_E_next = tot;
                *a.d_heavy_seen = static_cast<int64_t>(__ldcg(a.heavy));
            }
            const int cnt = static_cast<int>(t_hi - t_lo);
            constexpr int kPer = kReindexTilesPerBlock / 32;
            long long v[kPer], sum = 0;
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                const int x = lane * kPer + u;
                v[u] = x < cnt ? tsum[x] : 0;
                sum += v[u];
            }
            long long incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long long o = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += o;
            }
            long long run = base + incl - sum;
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                const int x = lane * kPer + u;
                if (x < cnt) a.tile_base_next[t_lo + x] = run;
                run += v[u];
            }
        }
    }
}
