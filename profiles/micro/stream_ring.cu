// Generator / tester ring in isolation: warp 0 advances 32 XORWOW generators and streams raw words through a shared-memory
// ring; warps 1..3 consume.  Variants isolate the cost of the named-barrier handshake and of the test work.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct X { uint32_t d, v0, v1, v2, v3, v4; };
__device__ __forceinline__ uint32_t raw(X &s) {
    const uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1; s.v1 = s.v2; s.v2 = s.v3; s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    return s.v4;
}
__device__ __forceinline__ void bsync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void barrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int CH = 32, NB = 4;
// MODE 0: handshake only (testers do nothing); 1: testers read the ring; 2: testers read + fastmod test against a table
template <int MODE>
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
#pragma unroll
            for (int t = 0; t < CH; t++) buf[p][t][lane] = raw(s);
            barrive(1 + p, 128);
        }
        const long long t1 = clock64();
        out[lane] = s.v4;
        if (lane == 0) *cyc = t1 - t0;
    } else {
        const int q = w - 1;
        uint32_t acc = 0;
        for (int p = 0; p < NB; p++) if (p < n_chunks) barrive(1 + NB + p, 128);
        for (int c = 0; c < n_chunks; c++) {
            const int p = c % NB;
            bsync(1 + p, 128);
            if (MODE >= 1) {
                uint32_t rr[11];
                unsigned long long mm[11];
#pragma unroll
                for (int u = 0; u < 11; u++) { const int t = q + 3 * u; rr[u] = t < CH ? buf[p][t][lane] : 0u; }
                if (MODE == 2) {
#pragma unroll
                    for (int u = 0; u < 11; u++) { const uint32_t idx = kk + lane + 32u * (c * CH + q + 3 * u); mm[u] = q + 3 * u < CH ? tab[idx] : 0ull; }
                    unsigned cand = 0;
#pragma unroll
                    for (int u = 0; u < 11; u++) cand |= (mm[u] * rr[u] < mm[u] * kk) ? 1u << u : 0u;
                    if (cand) {
#pragma unroll
                        for (int u = 0; u < 11; u++) if (cand >> u & 1u) {
                            const uint32_t idx = kk + lane + 32u * (c * CH + q + 3 * u);
                            const uint32_t num = (uint32_t)__umul64hi(mm[u] * rr[u], idx + 1);
                            if (num < kk) atomicMax(&slots[num], idx);
                        }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 11; u++) acc ^= rr[u];
                }
            }
            if (c + NB < n_chunks) barrive(1 + NB + p, 128);
        }
        out[32 + threadIdx.x] = acc + slots[lane & 7];
    }
}
__global__ void fill(unsigned long long *tab, uint32_t n) {
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x)
        tab[m] = m < 1 ? 0ull : (0xFFFFFFFFFFFFFFFFull / (m + 1) + 1ull);  // tab[idx] = recip[idx + 1]
}
template <int MODE>
void run(const char *name, const unsigned long long *tab) {
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int n_chunks = 90;
    for (int i = 0; i < 3; i++) ring<MODE><<<1, 128>>>(tab, out, cyc, n_chunks, 5);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %.1f cycles per round (%lld per chunk)\n", name, double(h) / (n_chunks * CH), h / n_chunks);
}
int main() {
    unsigned long long *tab; cudaMalloc(&tab, 8ull << 20);
    fill<<<256, 256>>>(tab, 1u << 20);
    run<0>("handshake only", tab);
    run<1>("testers read the ring", tab);
    run<2>("testers read + fastmod table test", tab);
    cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}
