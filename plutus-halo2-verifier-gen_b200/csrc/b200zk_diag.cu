// b200zk_diag.cu -- third translation unit of libb200zk.so: limb-for-limb self-test of the field layer and the
// integer-pipe micro-benchmarks that the roofline fractions of bench.py divide by (include/b200zk.h, last section).
// Runs on the calling thread's selected device (b200zk_set_device).  No CPU fallback.
#include <cuda_runtime.h>

#include "b200zk.h"
#include "ctx.hpp"
#include "field.cuh"
#include "g1.cuh"
#include "g1_call.cuh"

using namespace b200zk;
namespace ctx = b200zk_ctx;

namespace {

// ------------------------------------------------------------------------------------------
// self-test and micro-benchmark kernels
// ------------------------------------------------------------------------------------------
template <class P>
__global__ void selftest_kernel(const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t count, uint32_t op) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fe<P> x, y, r;
    for (int k = 0; k < P::N; k++) { x.l[k] = a[P::N * i + k]; y.l[k] = b ? b[P::N * i + k] : 0; }
    x = fe_to_mont(x);
    y = fe_to_mont(y);
    if (op == 0) r = fe_mul(x, y);
    else if (op == 1) r = fe_add(x, y);
    else if (op == 2) r = fe_sub(x, y);
    else r = fe_inv(x);
    r = fe_from_mont(r);
    for (int k = 0; k < P::N; k++) out[P::N * i + k] = r.l[k];
}

__global__ void __launch_bounds__(256) mb_imad_wide_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    // plain IMAD.WIDE.U32 (no carry in or out); the multiplicand rotates through the other
    // accumulators so that ptxas cannot hoist the product out of the loop
    uint64_t acc[8];
    uint32_t b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = a0 + threadIdx.x * 8 + k;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = (uint64_t)(uint32_t)acc[(k + 1) & 7] * b + acc[k];
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// carry-chained rows exactly as fe_mul issues them: 4 independent rows of 6 IMAD.WIDE.U32.X
__global__ void __launch_bounds__(256) mb_imad_chain_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint32_t acc[4][12];
    uint32_t a = a0 + threadIdx.x, b = b0 + blockIdx.x;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int k = 0; k < 12; k++) acc[r][k] = r * 12 + k;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            mad_wide_cc(acc[r][0], acc[r][1], a, b, acc[r][0], acc[r][1]);
#pragma unroll
            for (int k = 2; k < 12; k += 2) madc_wide_cc(acc[r][k], acc[r][k + 1], a, b, acc[r][k], acc[r][k + 1]);
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int k = 0; k < 12; k++) s = s * 31 + acc[r][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// every wide MAD produces a carry-out but takes no carry-in
__global__ void __launch_bounds__(256) mb_imad_cout_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint32_t lo[8], hi[8];
    uint32_t a = a0 + threadIdx.x, b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { lo[k] = k; hi[k] = k + 1; }
    uint32_t sink = 0;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) mad_wide_cc(lo[k], hi[k], a, b, lo[k], hi[k]);
        }
        sink = addc(sink, 0);
    }
    uint64_t s = sink;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= ((uint64_t)hi[k] << 32) | lo[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) mb_dfma_kernel(uint64_t* out, uint32_t iters, double a0, double b0) {
    double acc[8];
    double a = a0 + threadIdx.x * 1e-9, b = b0 + blockIdx.x * 1e-9;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = k;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[k]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)__double_as_longlong(s);
}
// one LOP3 (ALU pipe) per IMAD.WIDE (FMA pipe): do the two pipes issue side by side?
__global__ void __launch_bounds__(256) mb_imad_alu_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint64_t acc[8];
    uint32_t x[8];
    uint32_t b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = a0 + threadIdx.x * 8 + k; x[k] = k * 3; }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                acc[k] = (uint64_t)(uint32_t)acc[(k + 1) & 7] * b + acc[k];
                x[k] = (x[k] ^ x[(k + 3) & 7]) & ~x[(k + 5) & 7];
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k] + x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// MODE 0: IMAD.HI.U32 only; 1: 32-bit IMAD only; 2: unfused pair IMAD + IMAD.HI.U32 on the same operands with an
// immediate multiplier (what ptxas emits for the m*p rows when the modulus limb is not in a plain register).
// The multiplicand rotates through the accumulators so that nothing can be hoisted.
template <int MODE>
__global__ void __launch_bounds__(256) mb_imad_parts_kernel(uint64_t* out, uint32_t iters, uint32_t a0, uint32_t b0) {
    uint32_t lo[8], hi[8];
    uint32_t b = b0 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { lo[k] = a0 + threadIdx.x * 8 + k; hi[k] = k + 1; }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[k]) : "r"(hi[(k + 1) & 7]), "r"(b));
                else if (MODE == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(lo[(k + 1) & 7]), "r"(b));
                else {
                    uint32_t m = lo[(k + 1) & 7];
                    asm volatile("mad.lo.cc.u32 %0, %2, 0x53bda402, %0;\n\tmadc.hi.u32 %1, %2, 0x53bda402, %1;"
                                 : "+r"(lo[k]), "+r"(hi[k]) : "r"(m));
                }
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= ((uint64_t)hi[k] << 32) | lo[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class P>
__global__ void __launch_bounds__(256) mb_femul_kernel(uint32_t* out, uint32_t iters) {
    Fe<P> x = fe_one<P>(), y = fe_one<P>();
    x.l[0] += threadIdx.x;
    y.l[1] += blockIdx.x;
    for (uint32_t it = 0; it < iters; it++) {
        x = fe_mul(x, y);
        y = fe_mul(y, x);
    }
    uint32_t s = 0;
    for (int k = 0; k < P::N; k++) s ^= x.l[k] ^ y.l[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the arithmetic of one radix-2 DIF butterfly, (x, y) -> (x + y, (x - y) * w), four independent butterflies per thread like a
// stage of the NTT kernel, with no memory traffic at all: what the integer pipes deliver for the butterfly alone, at the
// occupancy the launch configuration allows (kind 13: 16 warps / SM like the NTT kernel, kind 14: 64 warps / SM)
__global__ void __launch_bounds__(256, 2) mb_butterfly_kernel(uint32_t* out, uint32_t iters) {
    Fr v[8], w = fe_one<FrParams>();
#pragma unroll
    for (int j = 0; j < 8; j++) { v[j] = fe_one<FrParams>(); v[j].l[0] += threadIdx.x + j; }
    w.l[1] += blockIdx.x;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 2; q >= 0; q--) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j & (1 << q)) continue;
                Fr x = v[j], y = v[j | (1 << q)];
                v[j] = fe_add(x, y);
                v[j | (1 << q)] = fe_mul(fe_sub(x, y), w);
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++)
        for (int k = 0; k < 8; k++) s ^= v[j].l[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(128) mb_madd_kernel(const uint32_t* gen_mont, uint32_t* out, uint32_t iters) {
    G1Affine q = g1a_ldg(gen_mont, 0);
    G1Xyzz acc;
    xyzz_from_affine(acc, q, false);
    xyzz_dbl(acc);
    for (uint32_t k = 0; k < (threadIdx.x & 7); k++) xyzz_dbl(acc);
    for (uint32_t it = 0; it < iters; it++) xyzz_add_mixed(acc, q, false);
    uint32_t s = 0;
    for (int k = 0; k < 12; k++) s ^= acc.x.l[k] ^ acc.zzz.l[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


struct Scratch {   // per call: these entry points are diagnostics, not hot paths
    void* p = nullptr;
    ~Scratch() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" {

int32_t b200zk_selftest_field(uint32_t field, uint32_t op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t count) {
    ctx::Dev* d = nullptr;
    XTRY(ctx::current(&d));
    if (field > 1 || op > 3 || !a || !out || (op < 3 && !b)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "bad selftest arguments");
    if (count == 0) return B200ZK_OK;
    ctx::DeviceScope ds(ctx::ordinal(d));
    std::lock_guard<std::mutex> lk(ctx::mutex(d));
    cudaStream_t s = ctx::stream(d);
    size_t esz = field ? 48 : 32, bytes = esz * count;
    Scratch sc;
    XCU(cudaMalloc(&sc.p, 3 * bytes + 64));
    uint8_t* da = reinterpret_cast<uint8_t*>(sc.p);
    uint8_t* db = da + bytes;
    uint8_t* dout = db + bytes;
    XCU(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, s));
    if (b) XCU(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, s));
    unsigned grid = (unsigned)((count + 127) / 128);
    if (field == 0)
        XLAUNCH(selftest_kernel<FrParams>, grid, 128, 0, s, (const uint32_t*)da, b ? (const uint32_t*)db : nullptr, (uint32_t*)dout, count, op);
    else
        XLAUNCH(selftest_kernel<FpParams>, grid, 128, 0, s, (const uint32_t*)da, b ? (const uint32_t*)db : nullptr, (uint32_t*)dout, count, op);
    XCU(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, s));
    XCU(cudaStreamSynchronize(s));
    return B200ZK_OK;
}

int32_t b200zk_microbench(uint32_t kind, uint32_t iters, double* out_ops_per_s, double* out_ms) {
    ctx::Dev* d = nullptr;
    XTRY(ctx::current(&d));
    if (kind > 14 || kind == 1 || !out_ops_per_s) return ctx::fail(B200ZK_ERR_INVALID_ARG, "bad microbench arguments");
    ctx::DeviceScope ds(ctx::ordinal(d));
    uint32_t* d_gen = nullptr;
    if (kind == 3) XTRY(ctx::generator_dev(d, &d_gen, ctx::stream(d)));
    std::lock_guard<std::mutex> lk(ctx::mutex(d));
    cudaStream_t s = ctx::stream(d);
    int sms = ctx::sm_count(d);
    unsigned threads = (kind == 3) ? 128 : 256;
    unsigned blocks = (unsigned)sms * ((kind == 3) ? 3 : ((kind == 2) ? 4 : ((kind == 12 || kind == 13) ? 2 : 8)));
    Scratch sc;
    XCU(cudaMalloc(&sc.p, (size_t)blocks * threads * 8 + 64));
    uint64_t* o64 = reinterpret_cast<uint64_t*>(sc.p);
    uint32_t* o32 = reinterpret_cast<uint32_t*>(sc.p);
    cudaEvent_t e0, e1;
    XCU(cudaEventCreate(&e0));
    XCU(cudaEventCreate(&e1));
    double per_thread = 0;
    for (int rep = 0; rep < 2; rep++) {  // rep 0 warms up
        XCU(cudaEventRecord(e0, s));
        switch (kind) {
            case 0:
                XLAUNCH(mb_imad_wide_kernel, blocks, threads, 0, s, o64, iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            case 2:
                XLAUNCH(mb_femul_kernel<FpParams>, blocks, threads, 0, s, o32, iters);
                per_thread = 2.0 * iters;
                break;
            case 3:
                XLAUNCH(mb_madd_kernel, blocks, threads, 0, s, (const uint32_t*)d_gen, o32, iters);
                per_thread = 1.0 * iters;
                break;
            case 4:
                XLAUNCH(mb_femul_kernel<FrParams>, blocks, threads, 0, s, o32, iters);
                per_thread = 2.0 * iters;
                break;
            case 5:
                XLAUNCH(mb_imad_chain_kernel, blocks, threads, 0, s, o64, iters, 12345u, 67890u);
                per_thread = 24.0 * iters;
                break;
            case 6:
                XLAUNCH(mb_dfma_kernel, blocks, threads, 0, s, o64, iters, 1.000001, 0.999999);
                per_thread = 32.0 * iters;
                break;
            case 7:
                XLAUNCH(mb_imad_cout_kernel, blocks, threads, 0, s, o64, iters, 12345u, 67890u);
                per_thread = 24.0 * iters;
                break;
            case 12:   // Fr products at 16 warps / SM (the occupancy of the NTT kernel) instead of 64
                XLAUNCH(mb_femul_kernel<FrParams>, blocks, threads, 0, s, o32, iters);
                per_thread = 2.0 * iters;
                break;
            case 13:
            case 14:   // butterflies per second
                XLAUNCH(mb_butterfly_kernel, blocks, threads, 0, s, o32, iters);
                per_thread = 12.0 * iters;
                break;
            case 9:
                XLAUNCH(mb_imad_parts_kernel<0>, blocks, threads, 0, s, o64, iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            case 10:
                XLAUNCH(mb_imad_parts_kernel<1>, blocks, threads, 0, s, o64, iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            case 11:
                XLAUNCH(mb_imad_parts_kernel<2>, blocks, threads, 0, s, o64, iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
            default:
                XLAUNCH(mb_imad_alu_kernel, blocks, threads, 0, s, o64, iters, 12345u, 67890u);
                per_thread = 32.0 * iters;
                break;
        }
        XCU(cudaEventRecord(e1, s));
        XCU(cudaEventSynchronize(e1));
    }
    float ms = 0;
    XCU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out_ops_per_s = per_thread * (double)blocks * threads / (ms * 1e-3);
    if (out_ms) *out_ms = ms;
    return B200ZK_OK;
}

}  // extern "C"
