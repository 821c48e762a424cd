#include <cstdio>
#include <cstdlib>
#include "qv_common.cuh"
#include "qv_xorwow.cuh"
using namespace qv;
int main()
{
    const uint32_t *mats = xorwow_jump_matrices_host();
    int bad = 0;
    for (uint64_t seed = 0; seed < 5; seed++) {
        for (uint32_t n : {0u, 1u, 2u, 5u, 31u, 256u, 1000u, 4460u, 65537u, 1000003u}) {
            Xorwow a = xorwow_seed(seed * 977 + 3), b = a;
            for (uint32_t t = 0; t < 7; t++) xorwow_next(a), xorwow_next(b);  // some offset first
            for (uint32_t t = 0; t < n; t++) xorwow_next(a);
            xorwow_jump(b, n, mats);
            for (int t = 0; t < 4; t++)
                if (xorwow_next(a) != xorwow_next(b)) bad++;
        }
    }
    printf("%s\n", bad ? "MISMATCH" : "jump ok");
    return bad != 0;
}
