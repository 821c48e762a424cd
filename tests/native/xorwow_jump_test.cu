#include <cstdio>
#include <cstdlib>
#include "qv_common.cuh"
#include "qv_xorwow.cuh"
using namespace qv;
int main()
{
    const uint32_t *mats = xorwow_jump_matrices_host();
    const uint32_t *tabs = xorwow_jump_tables_host();
    int bad = 0;
    for (uint64_t seed = 0; seed < 5; seed++) {
        for (uint64_t n : {0ull, 1ull, 2ull, 5ull, 31ull, 256ull, 1000ull, 4460ull, 65537ull, 1000003ull}) {
            Xorwow a = xorwow_seed(seed * 977 + 3), b = a;
            for (uint32_t t = 0; t < 7; t++) xorwow_next(a), xorwow_next(b);  // some offset first
            for (uint64_t t = 0; t < n; t++) xorwow_next(a);
            Xorwow c = b;
            xorwow_jump(b, n, mats);
            xorwow_jump_nib(c, n, tabs);  // the form the kernels use (4-bit lookup tables)
            for (int t = 0; t < 4; t++) {
                const uint32_t want = xorwow_next(a);
                if (want != xorwow_next(b)) bad++;
                if (want != xorwow_next(c)) bad++;
            }
        }
    }
    printf("%s\n", bad ? "MISMATCH" : "jump ok");
    return bad != 0;
}
