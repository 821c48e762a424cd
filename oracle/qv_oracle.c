/*
 * qv_oracle.c -- CPU restatement of torch-quiver's sampler + feature-gather hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product (torch-quiver_b200/) never links or imports this file.
 *
 * Parity status: PINNED.
 *   - XORWOW stream: pinned against NVIDIA's own curand_kernel.h compiled for the host
 *     (oracle/curand_probe.cpp -> tests/golden/xorwow_kat.json).
 *   - counts / deg<=k rows / reindex / gather: pinned against the reference CPU extension compiled
 *     unmodified from /root/reference (oracle/build_ref.py -> tests/golden/ref_cpu_*.json) and the
 *     known-answer mini fixture of SURVEY.md 8(c).
 *   - deg>k sampled ids: the reference's own tests hold NO stored vectors for them (only the structural
 *     check tests/cpp/test_quiver_cpu.cpp:32-51); they are pinned by the kernel source restated below
 *     line by line plus the real cuRAND stream.
 *
 * All citations are relative to /root/reference unless they name a CUDA toolkit header.
 * Plain C11, no dependencies.  Build: see oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * XORWOW (cuRAND device API generator; third-party arithmetic, not in /root/reference).
 * Restates /usr/local/cuda/include/curand_kernel.h (CUDA 12.9):
 *   state layout              :150-156   (d, v[5])
 *   curand()                  :863-874
 *   _curand_init_inplace      :800-825   (seed scrambling, then subsequence / offset skip-ahead)
 *   skip-ahead                :316-334 (vector x matrix over GF(2)), :720-737 (sequence spacing 2^67)
 * cuRAND ships precomputed powers of the one-step matrix (curand_precalc.h); we derive the same
 * matrices from the generator itself (:565-584 shows cuRAND doing exactly that when no table exists).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t d;
    uint32_t v[5];
} qo_xorwow;

#define QO_NW 5            /* words of linear state      */
#define QO_NB (32 * QO_NW) /* bits of linear state = 160 */

static uint32_t xorwow_step_linear(uint32_t v[QO_NW])
{
    /* curand_kernel.h:865-871 */
    uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1];
    v[1] = v[2];
    v[2] = v[3];
    v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
    return v[4];
}

QO_API uint32_t qo_xorwow_next(qo_xorwow *s)
{
    uint32_t r = xorwow_step_linear(s->v);
    s->d += 362437u; /* curand_kernel.h:872 */
    return r + s->d; /* :873 */
}

/* matrix[r] (QO_NW words) = image of basis bit r under the map; row-vector convention of
 * __curand_matvec_inplace (curand_kernel.h:316-334). */
typedef struct {
    uint32_t row[QO_NB][QO_NW];
} qo_mat;

static void vecmat(uint32_t v[QO_NW], const qo_mat *m)
{
    uint32_t r[QO_NW] = {0, 0, 0, 0, 0};
    for (int i = 0; i < QO_NW; i++)
        for (int j = 0; j < 32; j++)
            if (v[i] & (1u << j))
                for (int k = 0; k < QO_NW; k++) r[k] ^= m->row[i * 32 + j][k];
    memcpy(v, r, sizeof r);
}

static void matmat(qo_mat *out, const qo_mat *a, const qo_mat *b)
{
    qo_mat *tmp = (qo_mat *)malloc(sizeof(qo_mat));
    for (int r = 0; r < QO_NB; r++) {
        memcpy(tmp->row[r], a->row[r], sizeof tmp->row[r]);
        vecmat(tmp->row[r], b);
    }
    memcpy(out, tmp, sizeof *out);
    free(tmp);
}

static void one_step_matrix(qo_mat *m)
{
    /* curand_kernel.h:565-584 (__curand_generate_skipahead_matrix_xor) */
    for (int i = 0; i < QO_NB; i++) {
        uint32_t v[QO_NW] = {0, 0, 0, 0, 0};
        v[i / 32] = 1u << (i & 31);
        xorwow_step_linear(v);
        memcpy(m->row[i], v, sizeof v);
    }
}

/* seq_pow[i] = (one step)^(2^67 * 2^i): skipping 2^i subsequences.  Built once. */
#define QO_SEQ_BITS 40
static qo_mat *g_seq_pow = NULL; /* [QO_SEQ_BITS] */
static qo_mat *g_off_pow = NULL; /* [64]: (one step)^(2^i) */

static void build_tables(void)
{
    if (g_seq_pow) return;
    qo_mat *off = (qo_mat *)malloc(sizeof(qo_mat) * 64);
    one_step_matrix(&off[0]);
    for (int i = 1; i < 64; i++) matmat(&off[i], &off[i - 1], &off[i - 1]);
    qo_mat *seq = (qo_mat *)malloc(sizeof(qo_mat) * QO_SEQ_BITS);
    /* 2^67 steps = off[63] squared four more times (2^64,65,66,67). XORWOW_SEQUENCE_SPACING = 67,
     * curand_precalc.h:54 */
    qo_mat cur;
    memcpy(&cur, &off[63], sizeof cur);
    for (int i = 0; i < 4; i++) matmat(&cur, &cur, &cur);
    memcpy(&seq[0], &cur, sizeof cur);
    for (int i = 1; i < QO_SEQ_BITS; i++) matmat(&seq[i], &seq[i - 1], &seq[i - 1]);
    g_off_pow = off;
    g_seq_pow = seq;
}

QO_API void qo_xorwow_init(uint64_t seed, uint64_t subsequence, uint64_t offset, qo_xorwow *s)
{
    build_tables();
    /* curand_kernel.h:807-818 */
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
    /* :819 _skipahead_sequence_inplace -- d is untouched (2^67 * 362437 == 0 mod 2^32, :735) */
    for (int i = 0; i < QO_SEQ_BITS && (subsequence >> i); i++)
        if ((subsequence >> i) & 1) vecmat(s->v, &g_seq_pow[i]);
    /* :820 _skipahead_inplace, d += 362437 * (uint)offset (:717) */
    for (int i = 0; i < 64 && (offset >> i); i++)
        if ((offset >> i) & 1) vecmat(s->v, &g_off_pow[i]);
    s->d += 362437u * (uint32_t)offset;
}

/* Dump the 5x160 matrix that skips `nseq` whole subsequences (used to pin the product's own table). */
QO_API void qo_xorwow_seq_matrix(uint64_t nseq, uint32_t *out /* [160*5] */)
{
    build_tables();
    for (int r = 0; r < QO_NB; r++) {
        uint32_t v[QO_NW] = {0, 0, 0, 0, 0};
        v[r / 32] = 1u << (r & 31);
        for (int i = 0; i < QO_SEQ_BITS && (nseq >> i); i++)
            if ((nseq >> i) & 1) vecmat(v, &g_seq_pow[i]);
        memcpy(out + r * QO_NW, v, sizeof v);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Sampler, step 1: per-seed counts, exclusive scan, total.
 * srcs/cpp/src/quiver/cuda/quiver_sample.cu:157-169 (degree -> cap_by(k) -> exclusive_scan -> reduce),
 * srcs/cpp/include/quiver/quiver.cu.hpp:30-48 (get_adj_diff), functor.cu.hpp:4-17 (cap_by).
 * n_nodes = indptr.size(0)-1; k < 0 means "no cap" (quiver_sample.cu:161).
 * ---------------------------------------------------------------------------------------------- */
QO_API int64_t qo_sample_counts(const int64_t *indptr, int64_t n_nodes, int64_t n_edges, const int64_t *seeds,
                                int64_t S, int64_t k, int64_t *counts, int64_t *out_ptr)
{
    int64_t tot = 0;
    for (int64_t i = 0; i < S; i++) {
        int64_t v = seeds[i];
        int64_t end = (v + 1 < n_nodes) ? indptr[v + 1] : n_edges; /* quiver.cu.hpp:44-46 */
        int64_t deg = end - indptr[v];
        int64_t c = (k >= 0 && deg > k) ? k : deg;
        counts[i] = c;
        out_ptr[i] = tot;
        tot += c;
    }
    return tot;
}

/* ------------------------------------------------------------------------------------------------
 * Sampler, step 2: the GPU row-wise kernel, restated serially with its exact work decomposition.
 * srcs/cpp/include/quiver/cuda_random.cu.hpp:7-69 launched by quiver.cu.hpp:380-403 with
 * BLOCK_WARPS=4, TILE_SIZE=64, block (32,4), grid ceil(S/64), rand_seed literal 0.
 *   :17-19  tile geometry: block b covers output rows [64b, min(64b+64,S)); warp w takes rows 64b+w, +4, ...
 *   :21-23  per-thread generator: curand_init(rand_seed*gridDim.x + blockIdx.x, threadIdx.y*32+threadIdx.x, 0)
 *   :33-38  deg <= k: copy the row verbatim
 *   :41-57  deg  > k: slots[0..k) = 0..k-1; lane l visits idx = k+l, k+l+32, ... < deg, draws
 *           num = curand() % (idx+1) and, if num < k, slots[num] = max(slots[num], idx)
 *   :61-64  out[j] = indices[row_start + slots[j]]
 *   :67     the lane's generator state persists over the warp's (up to 16) rows
 * ---------------------------------------------------------------------------------------------- */
static void sample_neighbor_gpu_impl(uint64_t rand_seed, int64_t k, int64_t S, const int64_t *seeds,
                                     const int64_t *indptr, const int64_t *indices, const int64_t *out_ptr,
                                     int64_t *out, int64_t *pos_out)
{
    const int BLOCK_WARPS = 4, TILE = 64, WARP = 32;
    const int64_t grid = (S + TILE - 1) / TILE;
    int64_t *slots = (int64_t *)malloc(sizeof(int64_t) * (size_t)(k > 0 ? k : 1));
    for (int64_t b = 0; b < grid; b++) {
        const int64_t last_row = ((b + 1) * TILE < S) ? (b + 1) * TILE : S;
        for (int w = 0; w < BLOCK_WARPS; w++) {
            qo_xorwow rng[32];
            for (int l = 0; l < WARP; l++)
                qo_xorwow_init(rand_seed * (uint64_t)grid + (uint64_t)b, (uint64_t)(w * WARP + l), 0, &rng[l]);
            for (int64_t out_row = b * TILE + w; out_row < last_row; out_row += BLOCK_WARPS) {
                const int64_t row = seeds[out_row];
                const int64_t start = indptr[row];
                const int64_t deg = indptr[row + 1] - start;
                const int64_t o = out_ptr[out_row];
                if (deg <= k) {
                    for (int64_t j = 0; j < deg; j++) {
                        out[o + j] = indices[start + j];
                        if (pos_out) pos_out[o + j] = start + j;
                    }
                } else {
                    for (int64_t j = 0; j < k; j++) slots[j] = j;
                    for (int l = 0; l < WARP; l++) {
                        for (int64_t idx = k + l; idx < deg; idx += WARP) {
                            /* `const int num = curand(&rng) % (idx + 1)` with int idx: unsigned modulo */
                            const uint32_t num = qo_xorwow_next(&rng[l]) % (uint32_t)(idx + 1);
                            if ((int64_t)num < k && slots[num] < idx) slots[num] = idx;
                        }
                    }
                    for (int64_t j = 0; j < k; j++) {
                        out[o + j] = indices[start + slots[j]];
                        if (pos_out) pos_out[o + j] = start + slots[j];
                    }
                }
            }
        }
    }
    free(slots);
}

QO_API void qo_sample_neighbor_gpu(uint64_t rand_seed, int64_t k, int64_t S, const int64_t *seeds,
                                   const int64_t *indptr, const int64_t *indices, const int64_t *out_ptr,
                                   int64_t *out)
{
    sample_neighbor_gpu_impl(rand_seed, k, S, seeds, indptr, indices, out_ptr, out, NULL);
}

/* Same walk, also reporting WHERE each pick sits in the CSR: pos_out[e] = indptr[seed] + slot -- the position whose
 * edge id the reference's sample kernel would emit (quiver.cu.hpp:90-126 `begin_id`), had any caller kept it
 * (sage_sampler.py:143 returns an empty e_id).  Checks the product's opt-in e_id output. */
QO_API void qo_sample_neighbor_gpu_pos(uint64_t rand_seed, int64_t k, int64_t S, const int64_t *seeds,
                                       const int64_t *indptr, const int64_t *indices, const int64_t *out_ptr,
                                       int64_t *out, int64_t *pos_out)
{
    sample_neighbor_gpu_impl(rand_seed, k, S, seeds, indptr, indices, out_ptr, out, pos_out);
}

/* ------------------------------------------------------------------------------------------------
 * The same sampler with the generator chains of long rows CUT INTO SEGMENTS -- an executable statement of why the
 * product's "mega row" path (torch-quiver_b200/csrc/qv_sample.cu: mega_segments) is bit-identical to the walk above.
 * A lane's stream is a function of how many draws it has made: state(n) = curand_init(seed, subsequence, offset = n)
 * (curand_kernel.h:820, the offset skip-ahead).  So for a row with more than `mega_draws` draws per lane
 *   - the draws of lane l, t in [0, c_l), c_l = ceil((deg - k - l) / 32), are grouped into segments of `seg` draws;
 *   - segment j of lane l starts from state(n_prev_l + j*seg), where n_prev_l counts the lane's draws on the warp's
 *     earlier rows -- no state is carried from one segment to the next, segments may run in any order;
 *   - hits go to slots by max (commutative), slots start at 0 and an untouched slot j reads as j (what the kernel's
 *     zero-initialised global reservoir does);
 *   - after the row the lane continues from state(n_prev_l + c_l).
 * Segments are deliberately evaluated in REVERSE order here.
 * ---------------------------------------------------------------------------------------------- */
QO_API void qo_sample_neighbor_gpu_split(uint64_t rand_seed, int64_t k, int64_t S, const int64_t *seeds,
                                         const int64_t *indptr, const int64_t *indices, const int64_t *out_ptr,
                                         int64_t *out, int64_t mega_draws, int64_t seg)
{
    const int BLOCK_WARPS = 4, TILE = 64, WARP = 32;
    const int64_t grid = (S + TILE - 1) / TILE;
    int64_t *slots = (int64_t *)malloc(sizeof(int64_t) * (size_t)(k > 0 ? k : 1));
    for (int64_t b = 0; b < grid; b++) {
        const int64_t last_row = ((b + 1) * TILE < S) ? (b + 1) * TILE : S;
        const uint64_t seed = rand_seed * (uint64_t)grid + (uint64_t)b;
        for (int w = 0; w < BLOCK_WARPS; w++) {
            uint64_t made[32]; /* draws made so far by each lane's generator */
            qo_xorwow rng[32];
            for (int l = 0; l < WARP; l++) {
                made[l] = 0;
                qo_xorwow_init(seed, (uint64_t)(w * WARP + l), 0, &rng[l]);
            }
            for (int64_t out_row = b * TILE + w; out_row < last_row; out_row += BLOCK_WARPS) {
                const int64_t row = seeds[out_row];
                const int64_t start = indptr[row];
                const int64_t deg = indptr[row + 1] - start;
                const int64_t o = out_ptr[out_row];
                if (deg <= k) {
                    for (int64_t j = 0; j < deg; j++) out[o + j] = indices[start + j];
                    continue;
                }
                const int64_t c0 = (deg - k + WARP - 1) / WARP; /* lane 0 draws the most */
                if (c0 <= mega_draws) { /* ordinary row: walk it */
                    for (int64_t j = 0; j < k; j++) slots[j] = j;
                    for (int l = 0; l < WARP; l++)
                        for (int64_t idx = k + l; idx < deg; idx += WARP) {
                            const uint32_t num = qo_xorwow_next(&rng[l]) % (uint32_t)(idx + 1);
                            made[l]++;
                            if ((int64_t)num < k && slots[num] < idx) slots[num] = idx;
                        }
                    for (int64_t j = 0; j < k; j++) out[o + j] = indices[start + slots[j]];
                    continue;
                }
                /* mega row: independent segments, last one first */
                for (int64_t j = 0; j < k; j++) slots[j] = 0;
                const int64_t n_seg = (c0 + seg - 1) / seg;
                for (int64_t sj = n_seg - 1; sj >= 0; sj--)
                    for (int l = 0; l < WARP; l++) {
                        const int64_t c_l = deg > k + l ? (deg - k - l + WARP - 1) / WARP : 0;
                        const int64_t t0 = sj * seg, t1 = (t0 + seg < c_l) ? t0 + seg : c_l;
                        if (t0 >= t1) continue;
                        qo_xorwow g;
                        qo_xorwow_init(seed, (uint64_t)(w * WARP + l), made[l] + (uint64_t)t0, &g);
                        for (int64_t t = t0; t < t1; t++) {
                            const int64_t idx = k + l + WARP * t;
                            const uint32_t num = qo_xorwow_next(&g) % (uint32_t)(idx + 1);
                            if ((int64_t)num < k && slots[num] < idx) slots[num] = idx;
                        }
                    }
                for (int l = 0; l < WARP; l++) { /* the owner steps over the row */
                    const int64_t c_l = deg > k + l ? (deg - k - l + WARP - 1) / WARP : 0;
                    made[l] += (uint64_t)c_l;
                    qo_xorwow_init(seed, (uint64_t)(w * WARP + l), made[l], &rng[l]);
                }
                for (int64_t j = 0; j < k; j++) out[o + j] = indices[start + (slots[j] > j ? slots[j] : j)];
            }
        }
    }
    free(slots);
}

/* ------------------------------------------------------------------------------------------------
 * Reindex: frontier = unique(concat(inputs, outputs)) in first-occurrence order; col_idx[e] = local id of
 * outputs[e]; row_idx[e] = position of the seed that produced e.
 * GPU semantics (duplicate seeds are merged): quiver_sample.cu:18-63 (FillWithDuplicates: min index wins,
 * prefix over "is first occurrence"), :244-251 (rewrite outputs to local ids), :338-351 (row_idx fill).
 * CPU twin with identical results for unique seeds: srcs/cpp/src/quiver/quiver.cpp:40-84.
 * Serial open-addressing table, nothing clever.  Returns F = |frontier|.
 * ---------------------------------------------------------------------------------------------- */
static uint64_t mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

QO_API int64_t qo_reindex(const int64_t *inputs, int64_t S, const int64_t *outputs, int64_t tot,
                          const int64_t *counts, int64_t *frontier, int64_t *row_idx, int64_t *col_idx)
{
    const int64_t n = S + tot;
    uint64_t cap = 16;
    while (cap < (uint64_t)(2 * n + 2)) cap <<= 1;
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * cap);
    int64_t *vals = (int64_t *)malloc(sizeof(int64_t) * cap);
    uint8_t *used = (uint8_t *)calloc(cap, 1);
    int64_t F = 0;
    for (int64_t i = 0; i < n; i++) {
        const int64_t key = i < S ? inputs[i] : outputs[i - S];
        uint64_t p = mix64((uint64_t)key) & (cap - 1);
        while (used[p] && keys[p] != key) p = (p + 1) & (cap - 1);
        if (!used[p]) {
            used[p] = 1;
            keys[p] = key;
            vals[p] = F;
            frontier[F++] = key;
        }
        if (i >= S) col_idx[i - S] = vals[p];
    }
    int64_t e = 0;
    for (int64_t i = 0; i < S; i++)
        for (int64_t j = 0; j < counts[i]; j++) row_idx[e++] = i; /* quiver_sample.cu:341-351 */
    free(keys);
    free(vals);
    free(used);
    return F;
}

/* ------------------------------------------------------------------------------------------------
 * Feature gather ("collect"): res[i,:] = shard(idx_i)[idx_i - offset(shard), :], a pure byte copy.
 * srcs/cpp/include/quiver/shard_tensor.cu.hpp:7-18 (find: first s with idx < offsets[s+1]) and :19-61.
 * The reference leaves rows with idx < 0 or idx >= total rows uninitialised (:49, torch::empty at
 * quiver_feature.cu:270); the B200 build DEFINES them as all-zero rows, and so does this oracle.
 * shard_ptrs[s] points at row offsets[s]; shard_pitch[s] is that shard's row pitch in bytes
 * (the reference stores rows densely: pitch == row_bytes).
 * ---------------------------------------------------------------------------------------------- */
QO_API void qo_gather(const void *const *shard_ptrs, const int64_t *shard_pitch, const int64_t *offsets,
                      int n_shards, const int64_t *indices, const int64_t *feature_order, int64_t n,
                      int64_t row_bytes, void *out)
{
    char *dst = (char *)out;
    for (int64_t i = 0; i < n; i++) {
        int64_t idx = indices[i];
        int s = -1;
        if (idx >= 0 && feature_order) {
            /* srcs/python/quiver/feature.py:300-301: node_idx = feature_order[node_idx] */
            idx = (idx < offsets[n_shards]) ? feature_order[idx] : -1;
        }
        if (idx >= 0)
            for (int t = 1; t <= n_shards; t++)
                if (idx < offsets[t]) {
                    s = t - 1;
                    break;
                }
        if (s < 0) {
            memset(dst + i * row_bytes, 0, (size_t)row_bytes);
            continue;
        }
        const char *src = (const char *)shard_ptrs[s] + (idx - offsets[s]) * shard_pitch[s];
        memcpy(dst + i * row_bytes, src, (size_t)row_bytes);
    }
}

/* ------------------------------------------------------------------------------------------------
 * One hop of access-probability propagation (SURVEY 8(f) "next" row).
 * srcs/cpp/include/quiver/cuda_random.cu.hpp:71-104 (cal_next), fp32 arithmetic in the same order.
 * ---------------------------------------------------------------------------------------------- */
QO_API void qo_cal_next(const float *last_prob, float *cur_prob, int64_t N, int k, const int64_t *indptr,
                        const int64_t *indices)
{
    for (int64_t row = 0; row < N; row++) {
        const int64_t start = indptr[row];
        const int64_t deg = indptr[row + 1] - start;
        float acc = 1.0f;
        if (deg == 0) {
            cur_prob[row] = 0;
            continue;
        }
        for (int64_t i = start; i < start + deg; i++) {
            const int64_t u = indices[i];
            const int64_t udeg = indptr[u + 1] - indptr[u];
            float skip;
            if (udeg == 0)
                skip = 1;
            else if (udeg <= k)
                skip = 1 - last_prob[u];
            else
                skip = 1 - last_prob[u] + last_prob[u] * (udeg - k) / udeg;
            acc *= skip;
        }
        /* `1 - (1 - p) * acc`: nvcc (default -fmad=true, which the reference build uses: setup.py:67) contracts this
         * into a single fused multiply-add; fmaf() reproduces that rounding on the CPU */
        cur_prob[row] = fmaf(-(1 - last_prob[row]), acc, 1.0f);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Structural validator of a sampled layer (the reference's only sampler assertion).
 * tests/cpp/test_quiver_cpu.cpp:32-51: per-seed count, membership in the seed's row, no duplicate
 * *positions*, total length.  Returns 0 when valid, else 1 + index of the first offending seed.
 * Works for multigraph rows too (duplicate neighbour ids are matched position by position).
 * ---------------------------------------------------------------------------------------------- */
QO_API int64_t qo_validate_sample(const int64_t *indptr, const int64_t *indices, const int64_t *seeds, int64_t S,
                                  int64_t k, const int64_t *counts, const int64_t *out, int64_t out_len)
{
    int64_t pos = 0;
    for (int64_t i = 0; i < S; i++) {
        const int64_t start = indptr[seeds[i]];
        const int64_t deg = indptr[seeds[i] + 1] - start;
        const int64_t want = (k >= 0 && deg > k) ? k : deg;
        if (counts[i] != want) return 1 + i;
        if (pos + want > out_len) return 1 + i;
        uint8_t *taken = (uint8_t *)calloc((size_t)(deg > 0 ? deg : 1), 1);
        for (int64_t j = 0; j < want; j++) {
            const int64_t id = out[pos + j];
            int64_t hit = -1;
            for (int64_t p = 0; p < deg; p++)
                if (!taken[p] && indices[start + p] == id) {
                    hit = p;
                    break;
                }
            if (hit < 0) {
                free(taken);
                return 1 + i;
            }
            taken[hit] = 1;
        }
        free(taken);
        pos += want;
    }
    return pos == out_len ? 0 : 1 + S;
}
