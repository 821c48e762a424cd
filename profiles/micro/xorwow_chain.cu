// How fast can ONE warp advance its 32 XORWOW generators?  (cycles per draw round; nvcc -arch=sm_100a -O3)
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
template <int U, bool STORE>
__global__ void chain(uint32_t *out, long long *cyc, int rounds) {
    __shared__ uint32_t buf[32][32];
    X s{1u, threadIdx.x * 2654435761u + 1, 362436069u, 521288629u, 88675123u, 5783321u + threadIdx.x};
    const long long t0 = clock64();
    for (int r = 0; r < rounds; r += U) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t v = raw(s);
            if (STORE) buf[u & 31][threadIdx.x & 31] = v;
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = s.v4 + buf[3][threadIdx.x & 31];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int U, bool STORE>
void run(const char *name, int threads) {
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int rounds = 32 * 4096;
    chain<U, STORE><<<1, threads>>>(out, cyc, rounds);
    chain<U, STORE><<<1, threads>>>(out, cyc, rounds);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s threads %3d: %.2f cycles per round\n", name, threads, double(h) / rounds);
}
int main() {
    run<1, false>("unroll 1, no store", 32);
    run<5, false>("unroll 5, no store", 32);
    run<32, false>("unroll 32, no store", 32);
    run<32, true>("unroll 32, STS", 32);
    run<32, true>("unroll 32, STS", 128);
    run<8, true>("unroll 8, STS", 32);
    return 0;
}
