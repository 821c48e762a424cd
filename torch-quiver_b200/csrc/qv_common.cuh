// qv_common.cuh -- shared helpers for libquiver_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "quiver_b200.h"

#ifndef __CUDA_ARCH_LIST__
#define __CUDA_ARCH_LIST__ 1000
#endif

namespace qv
{
// ---- error reporting (reported to the caller; the reference prints and exit(1)s: common.hpp:18-26) ----
std::string &last_error_slot();
int fail(int code, const char *fmt, ...);
extern std::atomic<long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define QV_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            cudaGetLastError();                                                                               \
            return ::qv::fail(_e == cudaErrorMemoryAllocation ? QV_ERR_NOMEM : QV_ERR_CUDA, "%s failed: %s (%s:%d)", \
                              #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                             \
        }                                                                                                     \
    } while (0)

#define QV_CHECK_LAUNCH(name)                                                                                 \
    do {                                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                                  \
        if (_e != cudaSuccess)                                                                                \
            return ::qv::fail(QV_ERR_CUDA, "launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e),   \
                              __FILE__, __LINE__);                                                            \
        ::qv::count_launch();                                                                                 \
    } while (0)

#define QV_REQUIRE(cond, ...)                                                                                 \
    do {                                                                                                      \
        if (!(cond)) return ::qv::fail(QV_ERR_INVALID, __VA_ARGS__);                                          \
    } while (0)

#define QV_TRY(expr)                                                                                          \
    do {                                                                                                      \
        int _rc = (expr);                                                                                     \
        if (_rc != QV_OK) return _rc;                                                                         \
    } while (0)

// RAII "run on this device, then restore" (the reference leaves the device switched: quiver_feature.cu:169-175)
struct DeviceGuard {
    int prev = -1;
    bool active = false;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) {
            active = cudaSetDevice(device) == cudaSuccess;
        }
    }
    ~DeviceGuard()
    {
        if (active) cudaSetDevice(prev);
    }
};

int sm_count(int device);
// qv_gather.cu: validate + enqueue a gather; d_n (optional) is a device-resident row count capped by the host bound n
int gather_enqueue(const qv_shard_table *table, const int64_t *indices, const int64_t *feature_order, int64_t n,
                   const int64_t *d_n, int64_t row_bytes, void *out, int variant, cudaStream_t st);

// ---- device helpers ----
constexpr int kWarp = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming 16-byte accesses: gathered rows have no reuse, keep them out of L1 (guide: Guideline 13)
__device__ __forceinline__ int4 ld_stream_v4(const void *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(void *p, const int4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
// L2 eviction-priority variants: the gathered payload is touched once (evict_first), the small index / order arrays are
// re-read by every call (evict_last), so the streaming rows do not push them out of the 126 MB L2.
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int4 ld_stream_v4_hint(const void *p, unsigned long long pol)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_stream_v4_hint(void *p, const int4 &v, unsigned long long pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ long long ld_keep_s64(const void *p, unsigned long long pol)
{
    long long r;
    asm volatile("ld.global.nc.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
    return r;
}

__device__ __forceinline__ int2 ld_stream_v2(const void *p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v2(void *p, const int2 &v)
{
    asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

}  // namespace qv
