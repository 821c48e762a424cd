// qv_xorwow.cuh -- XORWOW generator state compatible with cuRAND's device API stream, without cuRAND.
//
// The reference sampler draws from curand_init(seed = rand_seed*gridDim.x + blockIdx.x, subsequence = thread id in
// the 128-thread block, offset 0) EVERY launch (include/quiver/cuda_random.cu.hpp:21-23); the sub-sequence
// skip-ahead (2^67 steps per sub-sequence) costs each thread a chain of 160x160 GF(2) matrix products and dominates
// small launches.  Here the 128 skip-ahead maps P_q = (step^(2^67))^q are built once on the host (qv_xorwow.cu),
// kept in device memory in a lane-coalesced layout, and applied by one kernel that fills a cache of ready-to-use
// generator states [block][6 words][128 threads].  Sampling kernels then start from a 24-byte load.
//
// Arithmetic follows the published XORWOW recurrence (Marsaglia) with cuRAND's Weyl constant and seeding:
// CUDA 12.9 curand_kernel.h:800-825 (seeding), :863-874 (step), :316-334 / :720-737 (skip-ahead semantics).
#pragma once
#include <cstdint>

namespace qv
{
constexpr int kXorwowWords = 5;
constexpr int kXorwowBits = 160;
constexpr int kRngBlockThreads = 128;  // the reference's block(32,4): quiver.cu.hpp:384-387
constexpr int kRngStateWords = 6;      // d, v[0..4]

struct Xorwow {
    uint32_t d, v0, v1, v2, v3, v4;
};

__host__ __device__ __forceinline__ uint32_t xorwow_next(Xorwow &s)
{
    const uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1;
    s.v1 = s.v2;
    s.v2 = s.v3;
    s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return s.v4 + s.d;
}

// State of curand_init(seed, 0, 0): the seed scrambling only.
__host__ __device__ __forceinline__ Xorwow xorwow_seed(uint64_t seed)
{
    const uint32_t s0 = static_cast<uint32_t>(seed) ^ 0xaad26b49u;
    const uint32_t s1 = static_cast<uint32_t>(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    Xorwow s;
    s.d = 6615241u + t1 + t0;
    s.v0 = 123456789u + t0;
    s.v1 = 362436069u ^ t0;
    s.v2 = 521288629u + t1;
    s.v3 = 88675123u ^ t1;
    s.v4 = 5783321u + t0;
    return s;
}

// Jump-ahead by an arbitrary number of draws.  The five xorshift words evolve linearly over GF(2): state(t+n) = A^n *
// state(t), and the Weyl word just adds n * 362437.  `mats` holds A^(2^i), i = 0..kJumpBits-1, as [i][bit r (160)][word
// (5)] (row r = image of basis bit r); the jump multiplies in the matrices of n's set bits.  Branch-free inner loop (an
// AND mask per state bit) so the 32 lanes of a warp, whose states differ, execute the same instructions with uniform
// matrix loads.  ~1.6 k instructions per set bit of n: worth it for chains of thousands of draws (hub rows), not below.
constexpr int kJumpBits = 40;  // any jump a warp's 16 rows can add up to (15 rows x 2^32 / 32 draws)

__host__ __device__ inline void xorwow_jump(Xorwow &s, uint64_t n, const uint32_t *__restrict__ mats)
{
    uint32_t v[kXorwowWords] = {s.v0, s.v1, s.v2, s.v3, s.v4};
    for (int i = 0; i < kJumpBits; i++) {
        if (!((n >> i) & 1ull)) continue;  // n is (nearly) warp-uniform: at most two values per row across the lanes
        const uint32_t *m = mats + static_cast<size_t>(i) * kXorwowBits * kXorwowWords;
        uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
        for (int w = 0; w < kXorwowWords; w++) {
            const uint32_t bits = v[w];
#pragma unroll 8
            for (int j = 0; j < 32; j++) {
                const uint32_t mask = 0u - ((bits >> j) & 1u);
                const uint32_t *row = m + (w * 32 + j) * kXorwowWords;
                r0 ^= row[0] & mask;
                r1 ^= row[1] & mask;
                r2 ^= row[2] & mask;
                r3 ^= row[3] & mask;
                r4 ^= row[4] & mask;
            }
        }
        v[0] = r0; v[1] = r1; v[2] = r2; v[3] = r3; v[4] = r4;
    }
    s.v0 = v[0]; s.v1 = v[1]; s.v2 = v[2]; s.v3 = v[3]; s.v4 = v[4];
    s.d += static_cast<uint32_t>(n) * 362437u;  // mod 2^32
}

// Host: A^(2^i), i < kJumpBits, in the layout xorwow_jump reads (kJumpBits * 160 * 5 words, process lifetime).
const uint32_t *xorwow_jump_matrices_host();

// The same jump with the matrices expanded into 4-bit lookup tables: tab[i][q][x] (8 words, 5 used) = image under A^(2^i)
// of the state whose q-th nibble is x and whose other bits are 0.  A matrix product is then 40 look-ups instead of 160
// masked row additions (~6x fewer instructions) -- what makes segments of a few hundred draws worth cutting.
// 40 x 40 x 16 x 32 B = 819 KB, L2-resident.
constexpr int kJumpNibbles = kXorwowBits / 4;  // 40
constexpr int kJumpEntryWords = 8;
constexpr size_t kJumpTableWords = size_t(kJumpBits) * kJumpNibbles * 16 * kJumpEntryWords;

__host__ __device__ inline void xorwow_jump_nib(Xorwow &s, uint64_t n, const uint32_t *__restrict__ tab)
{
    uint32_t v[kXorwowWords] = {s.v0, s.v1, s.v2, s.v3, s.v4};
    for (int i = 0; i < kJumpBits; i++) {
        if (!((n >> i) & 1ull)) continue;
        const uint32_t *t = tab + static_cast<size_t>(i) * kJumpNibbles * 16 * kJumpEntryWords;
        uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#pragma unroll
        for (int w = 0; w < kXorwowWords; w++) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t x = (v[w] >> (4 * q)) & 15u;
                const uint32_t *e = t + (static_cast<size_t>(w * 8 + q) * 16 + x) * kJumpEntryWords;
#ifdef __CUDA_ARCH__
                const uint4 a = *reinterpret_cast<const uint4 *>(e);
                r0 ^= a.x; r1 ^= a.y; r2 ^= a.z; r3 ^= a.w;
#else
                r0 ^= e[0]; r1 ^= e[1]; r2 ^= e[2]; r3 ^= e[3];
#endif
                r4 ^= e[4];
            }
        }
        v[0] = r0; v[1] = r1; v[2] = r2; v[3] = r3; v[4] = r4;
    }
    s.v0 = v[0]; s.v1 = v[1]; s.v2 = v[2]; s.v3 = v[3]; s.v4 = v[4];
    s.d += static_cast<uint32_t>(n) * 362437u;  // mod 2^32
}

// Host: the nibble tables (kJumpTableWords words, process lifetime).
const uint32_t *xorwow_jump_tables_host();

// Host: the 128 sub-sequence skip matrices in device layout [bit r (160)][word (5)][q (128)] (uint32).
// Returns a pointer to a process-lifetime host array of 160*5*128 words.
const uint32_t *xorwow_subseq_matrices_host();

// Fill generator states for the reference launch geometry of `rows` seeds (rows = d_rows ? *d_rows : rows_arg):
// grid = ceil(rows/64); block b < grid, thread q gets curand_init(rand_seed*grid + b, q, 0).
// Layout: states[((b * 6) + word) * 128 + q].  `n_blocks` = blocks to launch (>= grid; the rest exit).
int xorwow_fill_states(const uint32_t *matrices_dev, uint64_t rand_seed, int64_t rows_arg, const int64_t *d_rows,
                       int64_t n_blocks, uint32_t *states_dev, cudaStream_t stream);

}  // namespace qv
