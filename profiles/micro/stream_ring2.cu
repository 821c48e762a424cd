// Second iteration of the generator / tester ring microbenchmark: variants of the generator step (xor chain depth, unroll)
// and of the tester (table loads pipelined one chunk ahead, rolled vs unrolled).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct X { uint32_t d, v0, v1, v2, v3, v4; };
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t raw(X &s) {   // T first: the chain through v4 is SHL -> LOP3
    const uint32_t t = s.v0 ^ (s.v0 >> 2);
    const uint32_t T = t ^ (t << 1);
    s.v0 = s.v1; s.v1 = s.v2; s.v2 = s.v3; s.v3 = s.v4;
    s.v4 = xor3(s.v4, s.v4 << 4, T);
    return s.v4;
}
__device__ __forceinline__ void bsync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void barrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int CH = 32, NB = 4;
template <int GU, int MODE>
__global__ void ring(const unsigned long long *tab, uint32_t *out, long long *cyc, int n_chunks, uint32_t kk) {
    __shared__ uint32_t buf[NB][CH][32];
    __shared__ uint32_t slots[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < 32) slots[threadIdx.x] = 0;
    __syncthreads();
    if (w == 0) {
        X s{1u, lane * 2654435761u + 1, 362436069u, 521288629u, 88675123u, 5783321u + lane};
        const long long t0 = clock64();
        for (int c = 0; c < n_chunks; c++) {
            const int p = c % NB;
            bsync(1 + NB + p, 128);
#pragma unroll 1
            for (int t0_ = 0; t0_ < CH; t0_ += GU) {
#pragma unroll
                for (int t = 0; t < GU; t++) buf[p][t0_ + t][lane] = raw(s);
            }
            barrive(1 + p, 128);
        }
        const long long t1 = clock64();
        out[lane] = s.v4;
        if (lane == 0) *cyc = t1 - t0;
    } else {
        const int q = w - 1;
        for (int p = 0; p < NB; p++) if (p < n_chunks) barrive(1 + NB + p, 128);
        constexpr int PER = 11;
        unsigned long long mm[PER], nx[PER];
        const uint32_t first = kk + lane;
#pragma unroll
        for (int u = 0; u < PER; u++) mm[u] = tab[first + 32u * (q + 3 * u)];
        for (int c = 0; c < n_chunks; c++) {
            const int p = c % NB;
            if (MODE == 1) {
#pragma unroll
                for (int u = 0; u < PER; u++) nx[u] = tab[first + 32u * ((c + 1) * CH + q + 3 * u)];  // next chunk's reciprocals
            }
            bsync(1 + p, 128);
            if (MODE == 2) { if (c + NB < n_chunks) barrive(1 + NB + p, 128); continue; }
            uint32_t rr[PER];
#pragma unroll
            for (int u = 0; u < PER; u++) { const int t = q + 3 * u; rr[u] = t < CH ? buf[p][t][lane] : 0u; }
            if (MODE == 0) {
#pragma unroll
                for (int u = 0; u < PER; u++) mm[u] = q + 3 * u < CH ? tab[first + 32u * (c * CH + q + 3 * u)] : 0ull;
            }
            unsigned cand = 0;
#pragma unroll
            for (int u = 0; u < PER; u++) cand |= (q + 3 * u < CH && mm[u] * rr[u] < mm[u] * kk) ? 1u << u : 0u;
            if (cand) {
                for (int u = 0; u < PER; u++) if (cand >> u & 1u) {
                    const uint32_t idx = first + 32u * (c * CH + q + 3 * u);
                    const unsigned long long M = tab[idx];
                    const uint32_t r = buf[p][q + 3 * u][lane];
                    const uint32_t num = (uint32_t)__umul64hi(M * r, idx + 1);
                    if (num < kk) atomicMax(&slots[num], idx);
                }
            }
            if (c + NB < n_chunks) barrive(1 + NB + p, 128);
            if (MODE == 1) {
#pragma unroll
                for (int u = 0; u < PER; u++) mm[u] = nx[u];
            }
        }
        out[32 + threadIdx.x] = slots[lane & 7];
    }
}
__global__ void fill(unsigned long long *tab, uint32_t n) {
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x)
        tab[m] = m < 1 ? 0ull : (0xFFFFFFFFFFFFFFFFull / (m + 1) + 1ull);
}
template <int GU, int MODE>
void run(const char *name, const unsigned long long *tab) {
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int n_chunks = 90;
    for (int i = 0; i < 3; i++) ring<GU, MODE><<<1, 128>>>(tab, out, cyc, n_chunks, 5);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-56s %.1f cycles per round (%lld per chunk)\n", name, double(h) / (n_chunks * CH), h / n_chunks);
}
int main() {
    unsigned long long *tab; cudaMalloc(&tab, 8ull << 20);
    fill<<<256, 256>>>(tab, 1u << 20);
    run<32, 2>("gen unroll 32, testers idle (handshake only)", tab);
    run<8, 2>("gen unroll 8,  testers idle (handshake only)", tab);
    run<32, 0>("gen unroll 32, loads after the barrier", tab);
    run<8, 0>("gen unroll 8,  loads after the barrier", tab);
    run<32, 1>("gen unroll 32, next chunk's loads before the barrier", tab);
    run<8, 1>("gen unroll 8,  next chunk's loads before the barrier", tab);
    run<4, 1>("gen unroll 4,  next chunk's loads before the barrier", tab);
    cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}
