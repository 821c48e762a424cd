// curand_probe.cpp -- TEST INFRASTRUCTURE.  Runs NVIDIA's own cuRAND XORWOW device code on the host so the
// oracle's (and the product's) XORWOW restatement can be pinned against the real third-party arithmetic.
// curand_kernel.h guards its qualifiers with `#if !defined(QUALIFIERS)` and falls back to the *_host
// precalculated matrices off-device (curand_kernel.h:605-623, 1610-1620), so predefining QUALIFIERS is enough.
// Build (oracle/Makefile): g++ -O2 -I/usr/local/cuda/include curand_probe.cpp -o _ref/curand_probe
// Usage: curand_probe <n_draws> <seed> <subseq> [<seed> <subseq> ...]  -> JSON lines on stdout.
#define QUALIFIERS static inline
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv)
{
    if (argc < 4 || (argc % 2) != 0) {
        fprintf(stderr, "usage: %s n seed subseq [seed subseq ...]\n", argv[0]);
        return 2;
    }
    const int n = atoi(argv[1]);
    printf("[\n");
    for (int a = 2; a + 1 < argc; a += 2) {
        const unsigned long long seed = strtoull(argv[a], nullptr, 10);
        const unsigned long long subseq = strtoull(argv[a + 1], nullptr, 10);
        curandStateXORWOW_t st;
        curand_init(seed, subseq, 0, &st);
        printf(" {\"seed\": %llu, \"subseq\": %llu, \"state\": [%u, %u, %u, %u, %u, %u], \"draws\": [", seed, subseq,
               st.d, st.v[0], st.v[1], st.v[2], st.v[3], st.v[4]);
        for (int i = 0; i < n; i++) printf("%s%u", i ? ", " : "", curand(&st));
        printf("]}%s\n", a + 3 < argc ? "," : "");
    }
    printf("]\n");
    return 0;
}
