// qv_runtime.cu -- devices, peer access, memory tiers and CUDA IPC behind the C ABI (include/quiver_b200.h).
// Replaces the host-side plumbing of srcs/cpp/src/quiver/cuda/quiver_feature.cu (init_p2p :378-421,
// can_device_access_peer :422-428, append :145-206, share_ipc :335-350) and the quiverRegister macro
// (include/quiver/quiver.cu.hpp:16-26) -- with errors returned instead of exit(1), devices restored after use,
// and memory that can be freed.
#include <unistd.h>

#include <cstdlib>

#include <mutex>
#include <unordered_map>

#include "qv_common.cuh"

namespace qv
{
std::string &last_error_slot()
{
    static thread_local std::string s;
    return s;
}

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_slot() = buf;
    return code;
}

std::atomic<long long> g_launches{0};

int sm_count(int device)
{
    static int cache[64] = {0};
    if (device >= 0 && device < 64 && cache[device]) return cache[device];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    if (device >= 0 && device < 64) cache[device] = n;
    return n;
}
}  // namespace qv

using namespace qv;

namespace
{
std::mutex g_reg_mu;
// host registrations we made: base -> (bytes). Lets qv_host_unregister undo exactly what was registered and makes
// repeated registration of the same tensor idempotent.
struct Registration {
    size_t bytes;
    int refs;
    bool ours;  // false: the range was already pinned by someone else (torch pin_memory, ...) -- never unregistered here
};
std::unordered_map<void *, Registration> g_registered;
}  // namespace

extern "C" {

int qv_abi_version(void) { return QV_ABI_VERSION; }

const char *qv_last_error(void) { return last_error_slot().c_str(); }

int64_t qv_launch_count(void) { return g_launches.load(); }

int qv_device_count(int *count)
{
    QV_REQUIRE(count != nullptr, "qv_device_count: count is NULL");
    QV_CUDA(cudaGetDeviceCount(count));
    return QV_OK;
}

int qv_can_device_access_peer(int src, int dst, int *ok)
{
    QV_REQUIRE(ok != nullptr, "qv_can_device_access_peer: ok is NULL");
    *ok = 0;
    if (src == dst) {
        *ok = 1;
        return QV_OK;
    }
    int a = 0, b = 0;
    QV_CUDA(cudaDeviceCanAccessPeer(&a, src, dst));
    QV_CUDA(cudaDeviceCanAccessPeer(&b, dst, src));
    *ok = (a && b) ? 1 : 0;
    return QV_OK;
}

int qv_init_p2p(const int *devices, int n_devices, int *n_enabled)
{
    QV_REQUIRE(devices != nullptr || n_devices == 0, "qv_init_p2p: devices is NULL");
    int prev = 0;
    QV_CUDA(cudaGetDevice(&prev));
    int enabled = 0;
    int rc = QV_OK;
    for (int i = 0; i < n_devices && rc == QV_OK; i++) {
        for (int j = 0; j < n_devices && rc == QV_OK; j++) {
            if (i == j) continue;
            const int src = devices[i], dst = devices[j];
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, src, dst);
            if (e != cudaSuccess) {
                rc = fail(QV_ERR_CUDA, "cudaDeviceCanAccessPeer(%d,%d): %s", src, dst, cudaGetErrorString(e));
                break;
            }
            if (!can) continue;
            e = cudaSetDevice(src);
            if (e == cudaSuccess) e = cudaDeviceEnablePeerAccess(dst, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();  // torch (or an earlier call) got there first: fine
                e = cudaSuccess;
            }
            if (e != cudaSuccess) {
                cudaGetLastError();
                rc = fail(QV_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d->%d): %s", src, dst, cudaGetErrorString(e));
                break;
            }
            enabled++;
        }
    }
    cudaSetDevice(prev);
    if (n_enabled) *n_enabled = enabled;
    return rc;
}

int qv_malloc(int device, size_t bytes, void **dev_ptr)
{
    QV_REQUIRE(dev_ptr != nullptr, "qv_malloc: dev_ptr is NULL");
    *dev_ptr = nullptr;
    DeviceGuard g(device);
    // Blocks of >= 2 MiB are rounded UP to a multiple of 2 MiB.  Measured on 2 x B200 (profiles/peer_probe2.py,
    // profiles/r2_peer_alloc_granularity.txt): a peer GPU gathers random 1 KiB rows out of a 51.2e9-byte cudaMalloc block at
    // 105 GB/s and out of a (40 GiB + 1 KiB) block at 60 GB/s -- but out of an exact 40 GiB block, or the first block padded
    // to a 2 MiB multiple, at 745 GB/s.  The driver maps a block for peers with large pages only when its size is a whole
    // number of them; otherwise the reader's address translation misses on nearly every row.  Local reads do not care.
    static const size_t granule = getenv("QV_MALLOC_GRANULE") ? static_cast<size_t>(atoll(getenv("QV_MALLOC_GRANULE")))
                                                                : (size_t(2) << 20);
    if (granule > 1 && bytes >= (size_t(2) << 20)) bytes = (bytes + granule - 1) / granule * granule;
    QV_CUDA(cudaMalloc(dev_ptr, bytes ? bytes : 16));
    return QV_OK;
}

int qv_free(int device, void *dev_ptr)
{
    if (!dev_ptr) return QV_OK;
    DeviceGuard g(device);
    QV_CUDA(cudaFree(dev_ptr));
    return QV_OK;
}

int qv_upload_rows(int device, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes,
                   size_t rows)
{
    if (rows == 0 || row_bytes == 0) return QV_OK;
    QV_REQUIRE(dst && src, "qv_upload_rows: NULL pointer");
    QV_REQUIRE(dst_pitch >= row_bytes && src_pitch >= row_bytes, "qv_upload_rows: pitch smaller than row");
    DeviceGuard g(device);
    if (dst_pitch == row_bytes && src_pitch == row_bytes) {
        QV_CUDA(cudaMemcpy(dst, src, row_bytes * rows, cudaMemcpyHostToDevice));
    } else {
        QV_CUDA(cudaMemcpy2D(dst, dst_pitch, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice));
    }
    return QV_OK;
}

int qv_copy_rows_device(int device, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes,
                        size_t rows, qv_stream_t stream)
{
    if (rows == 0 || row_bytes == 0) return QV_OK;
    QV_REQUIRE(dst && src, "qv_copy_rows_device: NULL pointer");
    QV_REQUIRE(dst_pitch >= row_bytes && src_pitch >= row_bytes, "qv_copy_rows_device: pitch smaller than row");
    DeviceGuard g(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dst_pitch == row_bytes && src_pitch == row_bytes) {
        QV_CUDA(cudaMemcpyAsync(dst, src, row_bytes * rows, cudaMemcpyDeviceToDevice, st));
    } else {
        QV_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, rows, cudaMemcpyDeviceToDevice, st));
    }
    return QV_OK;
}

int qv_memset(int device, void *dst, int value, size_t bytes)
{
    if (!bytes) return QV_OK;
    DeviceGuard g(device);
    QV_CUDA(cudaMemset(dst, value, bytes));
    return QV_OK;
}

int qv_host_register(int device, void *host_ptr, size_t bytes, void **dev_ptr)
{
    QV_REQUIRE(host_ptr && dev_ptr, "qv_host_register: NULL pointer");
    *dev_ptr = nullptr;
    if (bytes == 0) bytes = 1;
    DeviceGuard g(device);
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registered.find(host_ptr);
    if (it != g_registered.end() && it->second.bytes >= bytes) {
        it->second.refs++;  // same tensor shared by several tables: one registration, counted
    } else {
        if (it != g_registered.end() && it->second.ours) {
            // same base, LONGER range than the one we hold: the old mapping does not cover it -- replace it under the
            // lock (the users of the shorter range keep a valid alias: a registration of [p, p+n) yields the same device
            // address for p whatever n is) and carry their references over
            cudaError_t e = cudaHostUnregister(host_ptr);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(QV_ERR_CUDA, "cudaHostUnregister(%p) before re-registering %zu bytes: %s", host_ptr, bytes,
                            cudaGetErrorString(e));
            }
        }
        const int carried = it != g_registered.end() ? it->second.refs : 0;
        // One registration for the whole range: the reference splits it into 1e9-byte pieces whose boundaries are
        // not page aligned (quiver.cu.hpp:19-25); a single cudaHostRegister has no such seams.
        cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
        if (e == cudaErrorHostMemoryAlreadyRegistered) {
            cudaGetLastError();  // pinned by someone else (e.g. torch pin_memory): usable, not ours to undo -- recorded
            g_registered[host_ptr] = Registration{bytes, carried + 1, false};  // as foreign so that it is never unregistered here
        } else if (e != cudaSuccess) {
            cudaGetLastError();
            if (it != g_registered.end()) g_registered.erase(host_ptr);
            return fail(QV_ERR_CUDA, "cudaHostRegister(%p, %zu): %s", host_ptr, bytes, cudaGetErrorString(e));
        } else {
            g_registered[host_ptr] = Registration{bytes, carried + 1, true};
        }
    }
    QV_CUDA(cudaHostGetDevicePointer(dev_ptr, host_ptr, 0));
    return QV_OK;
}

int qv_host_unregister(void *host_ptr)
{
    if (!host_ptr) return QV_OK;
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registered.find(host_ptr);
    if (it == g_registered.end()) return QV_OK;  // not ours (or already undone)
    if (--it->second.refs > 0) return QV_OK;
    const bool ours = it->second.ours;
    g_registered.erase(it);
    if (ours) QV_CUDA(cudaHostUnregister(host_ptr));
    return QV_OK;
}

int qv_ipc_get_handle(int device, void *dev_ptr, unsigned char handle[QV_IPC_HANDLE_BYTES])
{
    QV_REQUIRE(dev_ptr && handle, "qv_ipc_get_handle: NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == QV_IPC_HANDLE_BYTES, "CUDA IPC handle size changed");
    DeviceGuard g(device);
    cudaIpcMemHandle_t h;
    QV_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, sizeof h);
    return QV_OK;
}

int qv_ipc_open_handle(int device, const unsigned char handle[QV_IPC_HANDLE_BYTES], void **dev_ptr)
{
    QV_REQUIRE(handle && dev_ptr, "qv_ipc_open_handle: NULL pointer");
    *dev_ptr = nullptr;
    DeviceGuard g(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    QV_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QV_OK;
}

int qv_ipc_close_handle(int device, void *dev_ptr)
{
    if (!dev_ptr) return QV_OK;
    DeviceGuard g(device);
    QV_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return QV_OK;
}

}  // extern "C"
