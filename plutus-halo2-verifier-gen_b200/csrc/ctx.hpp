// ctx.hpp -- what the translation units of libb200zk.so share: the per-device contexts owned by
// b200zk.cu (one per GPU bound by b200zk_init_devices; a single process drives all of them), the
// error / launch macros and the device-binding rule of the entry points:
//   * host-buffer entry points fan out over every bound device (MSM, NTT batches) or run on the
//     calling thread's selected device (b200zk_set_device, default: the first bound device);
//   * "_dev" entry points run on the device that owns their pointers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include "b200zk.h"

namespace b200zk_ctx {
struct Dev;                                           // one bound GPU (defined in b200zk.cu)

int32_t fail(int32_t code, const std::string& msg);   // records the calling thread's error text
int32_t current(Dev** out);                           // the calling thread's selected device; NOT_INIT without one
int32_t for_pointer(const void* p, Dev** out);        // the bound device that owns device pointer p (null: current())
int32_t by_index(int index, Dev** out);               // position in the b200zk_init_devices list
int device_count();
int ordinal(Dev* d);                                  // CUDA device ordinal
int index(Dev* d);                                    // position in the b200zk_init_devices list
std::mutex& mutex(Dev* d);                            // guards the device's caches and shared scratch
cudaStream_t stream(Dev* d);                          // utility stream of host-buffer entry points outside MSM / NTT
int sm_count(Dev* d);
void count_launch();                                  // b200zk_launch_count()
int32_t generator_dev(Dev* d, uint32_t** out, cudaStream_t s);   // Montgomery affine generator of G1 in that device's HBM
void on_shutdown(void (*fn)());                       // called by b200zk_shutdown before the streams die
// cross-stream ordering of a device's shared scratch buffers: bracket every use
int32_t ws_enter(Dev* d, cudaStream_t s);
int32_t ws_leave(Dev* d, cudaStream_t s);

// makes `ordinal` the calling thread's CUDA device for the lifetime of the object, then restores the caller's
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int ordinal) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != ordinal) cudaSetDevice(ordinal);
        else prev = -1;
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
}  // namespace b200zk_ctx

#define XCU(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return b200zk_ctx::fail(e_ == cudaErrorMemoryAllocation ? B200ZK_ERR_OOM : B200ZK_ERR_CUDA, b_);     \
        }                                                                                               \
    } while (0)
#define XLAUNCH(kern, grid, block, smem, stream, ...)                                                   \
    do {                                                                                                \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                       \
        b200zk_ctx::count_launch();                                                                     \
        XCU(cudaGetLastError());                                                                        \
    } while (0)
#define XTRY(expr)                                                                                      \
    do {                                                                                                \
        int32_t rc_ = (expr);                                                                           \
        if (rc_ != B200ZK_OK) return rc_;                                                               \
    } while (0)
