// qv_xorwow.cu -- host construction of the XORWOW sub-sequence skip maps and the state-cache fill kernel.
// See qv_xorwow.cuh for the why.  No cuRAND headers or libraries are used by the product.
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "qv_common.cuh"
#include "qv_xorwow.cuh"

namespace qv
{
namespace
{
// 160x160 bit matrix over GF(2); row r (5 words) is the image of basis bit r (row-vector convention).
struct BitMat {
    uint32_t row[kXorwowBits][kXorwowWords];
};

void step_linear(uint32_t v[kXorwowWords])
{
    const uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1];
    v[1] = v[2];
    v[2] = v[3];
    v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}

void apply(uint32_t v[kXorwowWords], const BitMat &m)
{
    uint32_t r[kXorwowWords] = {0, 0, 0, 0, 0};
    for (int w = 0; w < kXorwowWords; w++) {
        uint32_t bits = v[w];
        while (bits) {
            const int j = __builtin_ctz(bits);
            bits &= bits - 1;
            const uint32_t *src = m.row[w * 32 + j];
            for (int k = 0; k < kXorwowWords; k++) r[k] ^= src[k];
        }
    }
    std::memcpy(v, r, sizeof r);
}

void multiply(BitMat &out, const BitMat &a, const BitMat &b)  // out = a then b
{
    BitMat tmp;
    for (int r = 0; r < kXorwowBits; r++) {
        std::memcpy(tmp.row[r], a.row[r], sizeof tmp.row[r]);
        apply(tmp.row[r], b);
    }
    out = tmp;
}

std::vector<uint32_t> build_matrices()
{
    auto one = std::make_unique<BitMat>();
    for (int r = 0; r < kXorwowBits; r++) {
        uint32_t v[kXorwowWords] = {0, 0, 0, 0, 0};
        v[r / 32] = 1u << (r & 31);
        step_linear(v);
        std::memcpy(one->row[r], v, sizeof v);
    }
    // A = step^(2^67): one sub-sequence (XORWOW_SEQUENCE_SPACING, curand_precalc.h:54)
    auto A = std::make_unique<BitMat>(*one);
    for (int i = 0; i < 67; i++) multiply(*A, *A, *A);
    // P_q = A^q, q = 0..127, stored [r][word][q]
    std::vector<uint32_t> out(size_t(kXorwowBits) * kXorwowWords * kRngBlockThreads);
    auto P = std::make_unique<BitMat>();
    for (int r = 0; r < kXorwowBits; r++)
        for (int k = 0; k < kXorwowWords; k++) P->row[r][k] = (r / 32 == k) ? (1u << (r & 31)) : 0u;
    for (int q = 0; q < kRngBlockThreads; q++) {
        if (q) multiply(*P, *P, *A);
        for (int r = 0; r < kXorwowBits; r++)
            for (int k = 0; k < kXorwowWords; k++)
                out[(size_t(r) * kXorwowWords + k) * kRngBlockThreads + q] = P->row[r][k];
    }
    return out;
}
}  // namespace

const uint32_t *xorwow_jump_matrices_host()
{
    static std::once_flag once;
    static std::vector<uint32_t> mats;
    std::call_once(once, [] {
        auto P = std::make_unique<BitMat>();
        for (int r = 0; r < kXorwowBits; r++) {
            uint32_t v[kXorwowWords] = {0, 0, 0, 0, 0};
            v[r / 32] = 1u << (r & 31);
            step_linear(v);
            std::memcpy(P->row[r], v, sizeof v);
        }
        mats.resize(size_t(kJumpBits) * kXorwowBits * kXorwowWords);
        for (int i = 0; i < kJumpBits; i++) {
            if (i) multiply(*P, *P, *P);  // A^(2^i)
            std::memcpy(mats.data() + size_t(i) * kXorwowBits * kXorwowWords, P->row, sizeof P->row);
        }
    });
    return mats.data();
}

const uint32_t *xorwow_jump_tables_host()
{
    static std::once_flag once;
    static std::vector<uint32_t> tabs;
    std::call_once(once, [] {
        const uint32_t *mats = xorwow_jump_matrices_host();
        tabs.assign(kJumpTableWords, 0u);
        for (int i = 0; i < kJumpBits; i++) {
            const uint32_t *m = mats + size_t(i) * kXorwowBits * kXorwowWords;
            for (int q = 0; q < kJumpNibbles; q++)
                for (int x = 0; x < 16; x++) {
                    uint32_t *e = tabs.data() + ((size_t(i) * kJumpNibbles + q) * 16 + x) * kJumpEntryWords;
                    for (int bit = 0; bit < 4; bit++)
                        if (x >> bit & 1)
                            for (int k = 0; k < kXorwowWords; k++) e[k] ^= m[size_t(q * 4 + bit) * kXorwowWords + k];
                }
        }
    });
    return tabs.data();
}

const uint32_t *xorwow_subseq_matrices_host()
{
    static std::once_flag once;
    static std::vector<uint32_t> mats;
    std::call_once(once, [] { mats = build_matrices(); });
    return mats.data();
}

// One thread per (block b, sub-sequence q): state = P_q * seed_state(rand_seed*grid + b).
__global__ void __launch_bounds__(kRngBlockThreads)
    xorwow_fill_states_kernel(const uint32_t *__restrict__ mats, uint64_t rand_seed, int64_t rows_arg,
                              const int64_t *__restrict__ d_rows, uint32_t *__restrict__ states)
{
    const int64_t rows = d_rows ? *d_rows : rows_arg;
    const int64_t grid = (rows + 63) / 64;
    const int64_t b = blockIdx.x;
    if (b >= grid) return;
    const int q = threadIdx.x;
    const Xorwow s = xorwow_seed(rand_seed * static_cast<uint64_t>(grid) + static_cast<uint64_t>(b));
    const uint32_t in[kXorwowWords] = {s.v0, s.v1, s.v2, s.v3, s.v4};
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#pragma unroll
    for (int w = 0; w < kXorwowWords; w++) {
#pragma unroll 8
        for (int j = 0; j < 32; j++) {
            // the seed state is block-uniform, so this branch never diverges
            if ((in[w] >> j) & 1u) {
                const uint32_t *m = mats + (size_t(w * 32 + j) * kXorwowWords) * kRngBlockThreads + q;
                r0 ^= m[0 * kRngBlockThreads];
                r1 ^= m[1 * kRngBlockThreads];
                r2 ^= m[2 * kRngBlockThreads];
                r3 ^= m[3 * kRngBlockThreads];
                r4 ^= m[4 * kRngBlockThreads];
            }
        }
    }
    uint32_t *o = states + size_t(b) * kRngStateWords * kRngBlockThreads + q;
    o[0 * kRngBlockThreads] = s.d;  // sub-sequence skips leave d unchanged (2^67 * 362437 == 0 mod 2^32)
    o[1 * kRngBlockThreads] = r0;
    o[2 * kRngBlockThreads] = r1;
    o[3 * kRngBlockThreads] = r2;
    o[4 * kRngBlockThreads] = r3;
    o[5 * kRngBlockThreads] = r4;
}

int xorwow_fill_states(const uint32_t *matrices_dev, uint64_t rand_seed, int64_t rows_arg, const int64_t *d_rows,
                       int64_t n_blocks, uint32_t *states_dev, cudaStream_t stream)
{
    if (n_blocks <= 0) return QV_OK;
    QV_REQUIRE(n_blocks < (int64_t(1) << 31), "xorwow_fill_states: too many blocks");
    xorwow_fill_states_kernel<<<static_cast<unsigned>(n_blocks), kRngBlockThreads, 0, stream>>>(
        matrices_dev, rand_seed, rows_arg, d_rows, states_dev);
    QV_CHECK_LAUNCH("xorwow_fill_states_kernel");
    return QV_OK;
}

}  // namespace qv
