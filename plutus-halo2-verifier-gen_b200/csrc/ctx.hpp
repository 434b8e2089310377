// ctx.hpp -- what the second translation unit (b200zk_ext.cu: polynomial-side Fr kernels, SRS
// generation, batched decompression) shares with the device context owned by b200zk.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include "b200zk.h"

namespace b200zk_ctx {
int32_t fail(int32_t code, const std::string& msg);   // records the calling thread's error text
std::mutex& mutex();                                  // the context mutex every entry point takes
int32_t need_init();
cudaStream_t stream();                                // the context stream (host-buffer entry points)
int sm_count();
void count_launch();                                  // b200zk_launch_count()
int32_t generator_dev(uint32_t** out, cudaStream_t s);// Montgomery affine generator of G1 in HBM
void on_shutdown(void (*fn)());                       // called by b200zk_shutdown before the stream dies
// cross-stream ordering of the process-wide workspaces: bracket every use of a shared scratch buffer
int32_t ws_enter(cudaStream_t s);
int32_t ws_leave(cudaStream_t s);
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
