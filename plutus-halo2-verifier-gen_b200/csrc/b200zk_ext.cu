// b200zk_ext.cu -- second translation unit of libb200zk.so: the C ABI of the components next to the
// MSM / NTT hot path: device-memory helpers for a shim that does not link the CUDA
// runtime, polynomial-side Fr kernels (poly.cuh), SRS generation and batched point decompression
// (srs.cuh).  Shares the device context of b200zk.cu through ctx.hpp.  No CPU fallback here either.
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <vector>

#include "b200zk.h"
#include "ctx.hpp"
#include "poly.cuh"
#include "srs.cuh"

using namespace b200zk;
namespace ctx = b200zk_ctx;

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int32_t ensure(size_t bytes) {
        if (bytes <= cap) return B200ZK_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        XCU(cudaMalloc(&p, want));
        cap = want;
        return B200ZK_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct GateProgram {
    uint32_t *program = nullptr, *consts = nullptr, *t_inv = nullptr;
    int32_t* rotations = nullptr;
    uint32_t n_instr = 0, n_consts = 0, n_rot = 0, n_columns = 0, log_ext = 0, log_period = 0;
};

struct Ext {   // per bound device
    int ordinal = -1;
    Buf scratch, totals, stage, small, colptr, omega_pows, status;
    size_t small_off = 0, colptr_off = 0;
    uint32_t* fixed_table = nullptr;   // 32 x 256 multiples of G
    std::map<uint64_t, GateProgram> programs;
    uint64_t next_program = 1;
};
Ext g_ext[16];
bool g_hooked = false;
std::mutex g_hook_mu;

void ext_shutdown() {
    for (Ext& x : g_ext) {
        if (x.ordinal < 0) continue;
        ctx::DeviceScope ds(x.ordinal);
        Buf* all[] = {&x.scratch, &x.totals, &x.stage, &x.small, &x.colptr, &x.omega_pows, &x.status};
        for (Buf* b : all) b->release();
        if (x.fixed_table) { cudaFree(x.fixed_table); x.fixed_table = nullptr; }
        for (auto& kv : x.programs) {
            cudaFree(kv.second.program);
            cudaFree(kv.second.consts);
            cudaFree(kv.second.rotations);
            if (kv.second.t_inv) cudaFree(kv.second.t_inv);
        }
        x.programs.clear();
        x.small_off = 0;
        x.colptr_off = 0;
        x.ordinal = -1;
    }
}

// One entry-point invocation: binds the device (the owner of `hint`, else the calling thread's selected device), makes it
// current, takes the device mutex (the scratch buffers and rings below are shared by every caller of that device) and
// restores the caller's CUDA device on exit.
struct Call {
    ctx::Dev* d = nullptr;
    Ext* x = nullptr;
    int prev = -1;
    std::unique_lock<std::mutex> lk;
    int32_t begin(const void* hint = nullptr) {
        XTRY(hint ? ctx::for_pointer(hint, &d) : ctx::current(&d));
        return begin_bound();
    }
    int32_t begin_bound() {
        {
            std::lock_guard<std::mutex> hl(g_hook_mu);
            if (!g_hooked) { ctx::on_shutdown(ext_shutdown); g_hooked = true; }
        }
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != ctx::ordinal(d)) XCU(cudaSetDevice(ctx::ordinal(d)));
        else prev = -1;
        lk = std::unique_lock<std::mutex>(ctx::mutex(d));
        x = &g_ext[ctx::index(d)];
        x->ordinal = ctx::ordinal(d);
        return B200ZK_OK;
    }
    // gate-program handles carry the index of the device they were created on
    int32_t begin_handle(uint64_t handle) {
        XTRY(ctx::by_index((int)(handle >> 48) - 1, &d));
        return begin_bound();
    }
    cudaStream_t stream() const { return ctx::stream(d); }
    ~Call() {
        if (lk.owns_lock()) lk.unlock();
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// `count` canonical Fr values from the host -> Montgomery form in a small device ring; stream-ordered
constexpr size_t SMALL_BYTES = 1 << 20;
int32_t upload_fr(Ext& x, const uint8_t* v, uint32_t count, cudaStream_t s, uint32_t** out) {
    size_t bytes = (size_t)count * 32;
    if (bytes > SMALL_BYTES / 2) return ctx::fail(B200ZK_ERR_INVALID_ARG, "too many host scalars in one call");
    XTRY(x.small.ensure(SMALL_BYTES));
    if (x.small_off + bytes > SMALL_BYTES) x.small_off = 0;
    uint32_t* d = reinterpret_cast<uint32_t*>(x.small.as<uint8_t>() + x.small_off);
    x.small_off += (bytes + 63) & ~(size_t)63;
    XCU(cudaMemcpyAsync(d, v, bytes, cudaMemcpyHostToDevice, s));
    XLAUNCH(fr_convert_kernel2, (count + 255) / 256, 256, 0, s, (const uint32_t*)d, d, (uint64_t)count, 1u);
    *out = d;
    return B200ZK_OK;
}

// brackets an entry point's use of the shared scratch buffers / scalar ring (ctx.hpp)
struct WsGuard {
    ctx::Dev* d = nullptr;
    cudaStream_t s;
    bool active = false;
    int32_t enter(ctx::Dev* dev, cudaStream_t st) {
        d = dev;
        s = st;
        int32_t rc = ctx::ws_enter(dev, st);
        active = rc == B200ZK_OK;
        return rc;
    }
    ~WsGuard() { if (active) ctx::ws_leave(d, s); }
};

bool misaligned(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) != 0;
}
inline cudaStream_t S(void* stream) { return reinterpret_cast<cudaStream_t>(stream); }
inline unsigned blocks(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

int32_t fixed_table(ctx::Dev* d, Ext& x, cudaStream_t s) {
    if (x.fixed_table) return B200ZK_OK;
    uint32_t* d_gen = nullptr;
    XTRY(ctx::generator_dev(d, &d_gen, s));
    XCU(cudaMalloc(&x.fixed_table, (size_t)FIXED_WINDOWS * 256 * 96));
    XLAUNCH(g1_fixed_table32_kernel, FIXED_WINDOWS * 256 / 128, 128, 0, s, (const uint32_t*)d_gen, x.fixed_table);
    return B200ZK_OK;
}

int32_t batch_invert(Ext& x, const uint32_t* in, uint32_t* out, uint64_t n, cudaStream_t s) {
    if (n == 0) return B200ZK_OK;
    XTRY(x.scratch.ensure(n * 32));
    XLAUNCH(fr_batch_invert_kernel, blocks(n, BINV_TILE), BINV_THREADS, 0, s, in, out, x.scratch.as<uint32_t>(), n);
    return B200ZK_OK;
}

}  // namespace

extern "C" {

// ---- device memory for shims that do not link the CUDA runtime ---------------------------------
int32_t b200zk_dev_alloc(void** out, size_t bytes) {
    Call call;
    XTRY(call.begin());
    Ext& x = *call.x;
    (void)x;
    if (!out) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null out pointer");
    XCU(cudaMalloc(out, bytes ? bytes : 16));
    return B200ZK_OK;
}
int32_t b200zk_dev_free(void* p) {
    Call call;
    XTRY(call.begin(p));
    Ext& x = *call.x;
    (void)x;
    if (p) XCU(cudaFree(p));
    return B200ZK_OK;
}
int32_t b200zk_dev_upload(void* d_dst, const void* src, size_t bytes) {
    Call call;
    XTRY(call.begin(d_dst));
    Ext& x = *call.x;
    (void)x;
    if ((!d_dst || !src) && bytes) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (bytes) {
        XCU(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, call.stream()));
        XCU(cudaStreamSynchronize(call.stream()));
    }
    return B200ZK_OK;
}
int32_t b200zk_dev_download(void* dst, const void* d_src, size_t bytes) {
    Call call;
    XTRY(call.begin(d_src));
    Ext& x = *call.x;
    (void)x;
    if ((!dst || !d_src) && bytes) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (bytes) {
        XCU(cudaDeviceSynchronize());   // the source may have been produced on a caller's stream
        XCU(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, call.stream()));
        XCU(cudaStreamSynchronize(call.stream()));
    }
    return B200ZK_OK;
}

// ---- batched G1 decompression --------------------------------------------------------------------
int32_t b200zk_g1_decompress_dev(const void* d_compressed, uint64_t n, void* d_out_mont, void* d_out_canon, void* d_status,
                                 void* stream) {
    Call call;
    XTRY(call.begin(d_compressed));
    Ext& x = *call.x;
    (void)x;
    if (n == 0) return B200ZK_OK;
    if (!d_compressed || !d_status || (!d_out_mont && !d_out_canon)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_out_mont, d_out_canon)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    XLAUNCH(g1_decompress_kernel, blocks(n, 128), 128, 0, S(stream), reinterpret_cast<const uint8_t*>(d_compressed), n,
            reinterpret_cast<uint32_t*>(d_out_canon), reinterpret_cast<uint32_t*>(d_out_mont), reinterpret_cast<uint32_t*>(d_status));
    return B200ZK_OK;
}

int32_t b200zk_g1_decompress_batch(const uint8_t* compressed, uint64_t n, uint8_t* out_affine, uint32_t* status) {
    Call call;
    XTRY(call.begin());
    Ext& x = *call.x;
    (void)x;
    if (n == 0) return B200ZK_OK;
    if (!compressed || !out_affine) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    cudaStream_t s = call.stream();
    XTRY(x.stage.ensure(n * (48 + 96) + 64));
    XTRY(x.status.ensure(n * 4));
    uint8_t* d_out = x.stage.as<uint8_t>();
    uint8_t* d_in = d_out + n * 96;
    XCU(cudaMemcpyAsync(d_in, compressed, n * 48, cudaMemcpyHostToDevice, s));
    XLAUNCH(g1_decompress_kernel, blocks(n, 128), 128, 0, s, (const uint8_t*)d_in, n, reinterpret_cast<uint32_t*>(d_out),
            (uint32_t*)nullptr, x.status.as<uint32_t>());
    XCU(cudaMemcpyAsync(out_affine, d_out, n * 96, cudaMemcpyDeviceToHost, s));
    std::vector<uint32_t> st(n);
    XCU(cudaMemcpyAsync(st.data(), x.status.p, n * 4, cudaMemcpyDeviceToHost, s));
    XCU(cudaStreamSynchronize(s));
    uint64_t first_bad = n;
    for (uint64_t i = 0; i < n; i++) {
        if (status) status[i] = st[i];
        if (st[i] && first_bad == n) first_bad = i;
    }
    if (first_bad != n) {
        static const char* why[] = {"", "not a compressed encoding", "bad infinity encoding", "x is not below p", "x is not on the curve"};
        return ctx::fail(B200ZK_ERR_BAD_POINT, "decompress: point " + std::to_string(first_bad) + ": " + why[st[first_bad] <= 4 ? st[first_bad] : 0]);
    }
    return B200ZK_OK;
}

// ---- fixed-base multiplication and SRS generation ------------------------------------------------
int32_t b200zk_g1_fixed_mul_dev(const void* d_scalars, uint32_t scalar_fmt, uint64_t n, void* d_out_mont, void* stream) {
    Call call;
    XTRY(call.begin(d_out_mont));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    if (n == 0) return B200ZK_OK;
    if (!d_scalars || !d_out_mont) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (scalar_fmt > B200ZK_FMT_MONT) return ctx::fail(B200ZK_ERR_INVALID_ARG, "unknown scalar format");
    if (misaligned(d_scalars, d_out_mont)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    XTRY(fixed_table(call.d, x, S(stream)));
    XLAUNCH(g1_fixed_mul_kernel, blocks(n, 128), 128, 0, S(stream), (const uint32_t*)x.fixed_table,
            reinterpret_cast<const uint32_t*>(d_scalars), scalar_fmt == B200ZK_FMT_MONT ? 1u : 0u, n, reinterpret_cast<uint32_t*>(d_out_mont));
    return B200ZK_OK;
}

int32_t b200zk_srs_generate_dev(const uint8_t s_bytes[32], uint32_t k, const uint8_t omega[32], void* d_g_mont,
                                void* d_g_lagrange_mont, void* stream) {
    Call call;
    XTRY(call.begin(d_g_mont ? d_g_mont : d_g_lagrange_mont));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    if (!s_bytes || (!d_g_mont && !d_g_lagrange_mont) || (d_g_lagrange_mont && !omega)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (k > 28) return ctx::fail(B200ZK_ERR_INVALID_ARG, "srs: k > 28 is not supported");
    if (misaligned(d_g_mont, d_g_lagrange_mont)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    cudaStream_t st = S(stream);
    const uint64_t n = (uint64_t)1 << k;
    XTRY(fixed_table(call.d, x, st));
    uint32_t* d_s = nullptr;
    XTRY(upload_fr(x, s_bytes, 1, st, &d_s));
    XTRY(x.stage.ensure(n * 32));
    uint32_t* sc = x.stage.as<uint32_t>();
    if (d_g_mont) {   // g[i] = s^i * G
        XLAUNCH(fr_geometric_kernel, blocks(n, 256 * 8), 256, 0, st, (const uint32_t*)d_s, (const uint32_t*)nullptr, sc, n);
        XLAUNCH(g1_fixed_mul_kernel, blocks(n, 128), 128, 0, st, (const uint32_t*)x.fixed_table, (const uint32_t*)sc, 1u, n,
                reinterpret_cast<uint32_t*>(d_g_mont));
    }
    if (d_g_lagrange_mont) {   // g_lagrange[i] = omega^i (s^n - 1) / (n (s - omega^i)) * G
        uint32_t* d_w = nullptr;
        XTRY(upload_fr(x, omega, 1, st, &d_w));
        XTRY(x.omega_pows.ensure(n * 32));
        XTRY(x.status.ensure(64));
        uint32_t* wp = x.omega_pows.as<uint32_t>();
        uint32_t* flag = x.status.as<uint32_t>();
        XCU(cudaMemsetAsync(flag, 0, 4, st));
        XLAUNCH(fr_geometric_kernel, blocks(n, 256 * 8), 256, 0, st, (const uint32_t*)d_w, (const uint32_t*)nullptr, wp, n);
        XLAUNCH(srs_lagrange_pre_kernel, blocks(n, 256), 256, 0, st, (const uint32_t*)wp, (const uint32_t*)d_s, sc, n, flag);
        XTRY(batch_invert(x, sc, sc, n, st));
        // c = (s^n - 1) / n on the host side of the stream: s^n by k squarings in a one-thread kernel would do,
        // but the geometric kernel already gives s^n as element n of the power series: compute it as 2 elements
        uint32_t* d_c = nullptr;   // [s^n - 1, n] -> c
        {
            XTRY(x.small.ensure(SMALL_BYTES));
            if (x.small_off + 256 > SMALL_BYTES) x.small_off = 0;
            d_c = reinterpret_cast<uint32_t*>(x.small.as<uint8_t>() + x.small_off);
            x.small_off += 256;
        }
        XLAUNCH(srs_lagrange_const_kernel, 1, 1, 0, st, (const uint32_t*)d_s, k, d_c);
        XLAUNCH(fr_pointwise_kernel, blocks(n, 256), 256, 0, st, 0u, (const uint32_t*)sc, (const uint32_t*)wp, (const uint32_t*)nullptr, sc, n);
        XLAUNCH(fr_pointwise_kernel, blocks(n, 256), 256, 0, st, 3u, (const uint32_t*)sc, (const uint32_t*)nullptr, (const uint32_t*)d_c, sc, n);
        XLAUNCH(g1_fixed_mul_kernel, blocks(n, 128), 128, 0, st, (const uint32_t*)x.fixed_table, (const uint32_t*)sc, 1u, n,
                reinterpret_cast<uint32_t*>(d_g_lagrange_mont));
        uint32_t bad = 0;
        XCU(cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, st));
        XCU(cudaStreamSynchronize(st));
        if (bad) return ctx::fail(B200ZK_ERR_INVALID_ARG, "srs: the secret lies in the evaluation domain (s = omega^i)");
    }
    return B200ZK_OK;
}

int32_t b200zk_g1_export_dev(const void* d_points_mont, uint64_t n, uint8_t* out_affine) {
    Call call;
    XTRY(call.begin(d_points_mont));
    Ext& x = *call.x;
    (void)x;
    if (n == 0) return B200ZK_OK;
    if (!d_points_mont || !out_affine) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    cudaStream_t s = call.stream();
    XCU(cudaDeviceSynchronize());
    XTRY(x.scratch.ensure(n * 96));
    XLAUNCH(g1_to_canonical_kernel, blocks(n, 128), 128, 0, s, reinterpret_cast<const uint32_t*>(d_points_mont), x.scratch.as<uint32_t>(), n);
    XCU(cudaMemcpyAsync(out_affine, x.scratch.p, n * 96, cudaMemcpyDeviceToHost, s));
    XCU(cudaStreamSynchronize(s));
    return B200ZK_OK;
}

// ---- Fr vectors ----------------------------------------------------------------------------------
int32_t b200zk_fr_convert_dev(const void* d_in, void* d_out, uint64_t n, uint32_t to_mont, void* stream) {
    Call call;
    XTRY(call.begin(d_out));
    Ext& x = *call.x;
    (void)x;
    if (n == 0) return B200ZK_OK;
    if (!d_in || !d_out) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_in, d_out)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    XLAUNCH(fr_convert_kernel2, blocks(n, 256), 256, 0, S(stream), reinterpret_cast<const uint32_t*>(d_in),
            reinterpret_cast<uint32_t*>(d_out), n, to_mont ? 1u : 0u);
    return B200ZK_OK;
}

int32_t b200zk_fr_power_table_dev(const uint8_t base[32], uint64_t row0, uint64_t rows, uint64_t cols, void* d_out, void* stream) {
    Call call;
    XTRY(call.begin(d_out));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    if (rows == 0 || cols == 0) return B200ZK_OK;
    if (!base || !d_out) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_out)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    if (row0 + rows > (1ull << 32) || cols > (1ull << 32)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "power table: exponents exceed 64 bits");
    uint32_t* d_b = nullptr;
    XTRY(upload_fr(x, base, 1, S(stream), &d_b));
    XLAUNCH(fr_power_table_kernel, blocks(rows * cols, 256), 256, 0, S(stream), (const uint32_t*)d_b, row0, rows, cols,
            reinterpret_cast<uint32_t*>(d_out));
    return B200ZK_OK;
}

int32_t b200zk_fr_extend_dev(const void* d_in, uint64_t n_in, void* d_out, uint64_t n_out, uint32_t batch, void* stream) {
    Call call;
    XTRY(call.begin(d_out));
    Ext& x = *call.x;
    (void)x;
    if (batch == 0 || n_out == 0) return B200ZK_OK;
    if (!d_in || !d_out) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (n_in > n_out) return ctx::fail(B200ZK_ERR_INVALID_ARG, "extend: n_in exceeds n_out");
    if (misaligned(d_in, d_out)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    // every polynomial: n_in coefficients followed by zeros up to n_out (the zero padding of coeff_to_extended)
    XCU(cudaMemsetAsync(d_out, 0, (size_t)batch * n_out * 32, S(stream)));
    if (n_in) XCU(cudaMemcpy2DAsync(d_out, n_out * 32, d_in, n_in * 32, n_in * 32, batch, cudaMemcpyDeviceToDevice, S(stream)));
    return B200ZK_OK;
}

int32_t b200zk_fr_pointwise_dev(uint32_t op, const void* d_a, const void* d_b, const uint8_t scalar[32], void* d_out, uint64_t n,
                                void* stream) {
    Call call;
    XTRY(call.begin(d_out));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    if (op > 4) return ctx::fail(B200ZK_ERR_INVALID_ARG, "pointwise: unknown op");
    if (n == 0) return B200ZK_OK;
    if (!d_a || !d_out || (op != 3 && !d_b) || (op == 3 && !scalar)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_a, d_b, d_out)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    uint32_t* d_s = nullptr;
    if (op == 3) XTRY(upload_fr(x, scalar, 1, S(stream), &d_s));
    XLAUNCH(fr_pointwise_kernel, blocks(n, 256), 256, 0, S(stream), op, reinterpret_cast<const uint32_t*>(d_a),
            reinterpret_cast<const uint32_t*>(d_b), (const uint32_t*)d_s, reinterpret_cast<uint32_t*>(d_out), n);
    return B200ZK_OK;
}

int32_t b200zk_fr_lincomb_dev(const void* const* d_polys, const uint8_t* coeffs, uint32_t count, void* d_out, uint64_t n, void* stream) {
    Call call;
    XTRY(call.begin(d_out));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    if (!d_out && n) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (n == 0) return B200ZK_OK;
    if (count == 0) { XCU(cudaMemsetAsync(d_out, 0, n * 32, S(stream))); return B200ZK_OK; }
    if (!d_polys || !coeffs) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_out)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    for (uint32_t done = 0; done < count; done += LINCOMB_MAX) {
        LincombArgs a;
        memset(&a, 0, sizeof a);
        a.count = count - done < (uint32_t)LINCOMB_MAX ? count - done : (uint32_t)LINCOMB_MAX;
        a.accumulate = done ? 1u : 0u;
        for (uint32_t j = 0; j < a.count; j++) {
            if (!d_polys[done + j] || misaligned(d_polys[done + j])) return ctx::fail(B200ZK_ERR_INVALID_ARG, "lincomb: null or misaligned polynomial");
            a.poly[j] = reinterpret_cast<const uint32_t*>(d_polys[done + j]);
        }
        uint32_t* d_c = nullptr;
        XTRY(upload_fr(x, coeffs + 32 * (size_t)done, a.count, S(stream), &d_c));
        XLAUNCH(fr_lincomb_kernel, blocks(n, 256), 256, 0, S(stream), a, (const uint32_t*)d_c, reinterpret_cast<uint32_t*>(d_out), n);
    }
    return B200ZK_OK;
}

int32_t b200zk_fr_batch_invert_dev(const void* d_in, void* d_out, uint64_t n, void* stream) {
    Call call;
    XTRY(call.begin(d_out));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    if (n == 0) return B200ZK_OK;
    if (!d_in || !d_out) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_in, d_out)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    return batch_invert(x, reinterpret_cast<const uint32_t*>(d_in), reinterpret_cast<uint32_t*>(d_out), n, S(stream));
}

int32_t b200zk_fr_running_product_dev(const void* d_in, void* d_out, uint64_t n, const uint8_t init[32], uint32_t inclusive,
                                      void* stream) {
    Call call;
    XTRY(call.begin(d_out));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    if (n == 0) return B200ZK_OK;
    if (!d_in || !d_out) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_in, d_out)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    cudaStream_t s = S(stream);
    uint64_t tiles = (n + PS_TILE - 1) / PS_TILE;
    XTRY(x.totals.ensure(tiles * 64));
    uint32_t* tot = x.totals.as<uint32_t>();
    uint32_t* d_init = nullptr;
    if (init) XTRY(upload_fr(x, init, 1, s, &d_init));
    const uint32_t* in = reinterpret_cast<const uint32_t*>(d_in);
    XLAUNCH(fr_prodscan_totals_kernel, (unsigned)tiles, PS_THREADS, 0, s, in, n, tot);
    XLAUNCH(fr_prodscan_middle_kernel, 1, 256, 0, s, tot, tiles, (const uint32_t*)d_init);
    XLAUNCH(fr_prodscan_apply_kernel, (unsigned)tiles, PS_THREADS, 0, s, in, reinterpret_cast<uint32_t*>(d_out), n, (const uint32_t*)tot,
            inclusive ? 1u : 0u);
    return B200ZK_OK;
}

int32_t b200zk_fr_kate_div_dev(const void* d_coeffs, uint64_t n, const uint8_t z[32], void* d_quot, void* d_eval, void* stream) {
    Call call;
    XTRY(call.begin(d_quot ? d_quot : d_eval));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    if (!z || (!d_coeffs && n) || (!d_quot && !d_eval)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_coeffs, d_quot, d_eval)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    cudaStream_t s = S(stream);
    if (n == 0) { if (d_eval) XCU(cudaMemsetAsync(d_eval, 0, 32, s)); return B200ZK_OK; }
    uint64_t tiles = (n + PS_TILE - 1) / PS_TILE;
    XTRY(x.totals.ensure(tiles * 64));
    uint32_t* tot = x.totals.as<uint32_t>();
    // zpow = {z, z^PS_ITEMS}: upload z twice and raise the second copy on the device
    uint8_t zz[64];
    memcpy(zz, z, 32);
    memcpy(zz + 32, z, 32);
    uint32_t* d_z = nullptr;
    XTRY(upload_fr(x, zz, 2, s, &d_z));
    XLAUNCH(fr_pow_small_kernel, 1, 1, 0, s, d_z + 8, (uint32_t)PS_ITEMS);
    const uint32_t* c = reinterpret_cast<const uint32_t*>(d_coeffs);
    XLAUNCH(fr_horner_totals_kernel, (unsigned)tiles, PS_THREADS, 0, s, c, n, (const uint32_t*)d_z, tot);
    XLAUNCH(fr_horner_middle_kernel, 1, 256, 0, s, tot, tiles);
    XLAUNCH(fr_horner_apply_kernel, (unsigned)tiles, PS_THREADS, 0, s, c, n, (const uint32_t*)d_z, (const uint32_t*)tot,
            reinterpret_cast<uint32_t*>(d_quot), reinterpret_cast<uint32_t*>(d_eval));
    return B200ZK_OK;
}

// ---- gate programs -------------------------------------------------------------------------------
int32_t b200zk_gate_program_create(const uint32_t* program, uint32_t n_instr, const uint8_t* consts, uint32_t n_consts,
                                   const int32_t* rotations, uint32_t n_rotations, const uint8_t* t_inv, uint32_t log_period,
                                   uint32_t n_columns, uint32_t log_n, uint32_t log_ext, uint64_t* out_handle) {
    Call call;
    XTRY(call.begin());
    Ext& x = *call.x;
    (void)x;
    if (!out_handle || (!program && n_instr) || (!consts && n_consts) || (!rotations && n_rotations))
        return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (log_ext < log_n || log_ext > 32 || n_rotations > 0x1000 || n_columns > 0x10000 || n_consts >= (1u << 28) || log_period > log_ext)
        return ctx::fail(B200ZK_ERR_INVALID_ARG, "gate program: domain, rotation, column or constant count out of range");
    // validate: every source names something that exists, registers are written before they are read
    std::vector<uint8_t> written(GATE_MAX_REGS, 0);
    for (uint32_t pc = 0; pc < n_instr; pc++) {
        const uint32_t* ins = program + 4 * (size_t)pc;
        uint32_t op = ins[0] & 0xffu, dst = ins[0] >> 8;
        if (op > 7 || dst >= (uint32_t)GATE_MAX_REGS)
            return ctx::fail(B200ZK_ERR_INVALID_ARG, "gate program: instruction " + std::to_string(pc) + ": bad opcode or destination register");
        int nsrc = (op == 3 || op == 4 || op == 5 || op == 7) ? 1 : (op == 6 ? 3 : 2);
        for (int k = 0; k < nsrc; k++) {
            uint32_t src = ins[1 + k], kind = src >> 28, pay = src & 0x0fffffffu;
            bool ok = (kind == 0 && pay < n_consts) || (kind == 1 && pay < (uint32_t)GATE_MAX_REGS && written[pay]) ||
                      (kind == 2 && (pay >> 12) < n_columns && (pay & 0xfffu) < n_rotations);
            if (!ok) return ctx::fail(B200ZK_ERR_INVALID_ARG, "gate program: instruction " + std::to_string(pc) + ": bad source operand " + std::to_string(k));
        }
        written[dst] = 1;
    }
    cudaStream_t s = call.stream();
    GateProgram gp;
    gp.n_instr = n_instr; gp.n_consts = n_consts; gp.n_rot = n_rotations; gp.n_columns = n_columns;
    gp.log_ext = log_ext; gp.log_period = log_period;
    XCU(cudaMalloc(&gp.program, (size_t)(n_instr ? n_instr : 1) * 16));
    XCU(cudaMalloc(&gp.consts, (size_t)(n_consts ? n_consts : 1) * 32));
    XCU(cudaMalloc(&gp.rotations, (size_t)(n_rotations ? n_rotations : 1) * 4));
    if (n_instr) XCU(cudaMemcpyAsync(gp.program, program, (size_t)n_instr * 16, cudaMemcpyHostToDevice, s));
    if (n_consts) {
        XCU(cudaMemcpyAsync(gp.consts, consts, (size_t)n_consts * 32, cudaMemcpyHostToDevice, s));
        XLAUNCH(fr_convert_kernel2, blocks(n_consts, 256), 256, 0, s, (const uint32_t*)gp.consts, gp.consts, (uint64_t)n_consts, 1u);
    }
    std::vector<int32_t> rot(n_rotations);
    for (uint32_t i = 0; i < n_rotations; i++) {
        // scaled to the extended domain in 64 bits and reduced mod 2^log_ext (a rotation is an index offset on a cyclic domain)
        int64_t r = (int64_t)rotations[i] * ((int64_t)1 << (log_ext - log_n));
        int64_t m = (int64_t)1 << log_ext;
        r %= m;
        if (r >= m / 2) r -= m;
        if (r < -(m / 2)) r += m;
        rot[i] = (int32_t)r;
    }
    if (n_rotations) XCU(cudaMemcpyAsync(gp.rotations, rot.data(), (size_t)n_rotations * 4, cudaMemcpyHostToDevice, s));
    if (t_inv) {
        uint64_t cnt = (uint64_t)1 << log_period;
        XCU(cudaMalloc(&gp.t_inv, cnt * 32));
        XCU(cudaMemcpyAsync(gp.t_inv, t_inv, cnt * 32, cudaMemcpyHostToDevice, s));
        XLAUNCH(fr_convert_kernel2, blocks(cnt, 256), 256, 0, s, (const uint32_t*)gp.t_inv, gp.t_inv, cnt, 1u);
    }
    XCU(cudaStreamSynchronize(s));
    uint64_t h = ((uint64_t)(ctx::index(call.d) + 1) << 48) | x.next_program++;
    x.programs[h] = gp;
    *out_handle = h;
    return B200ZK_OK;
}

int32_t b200zk_gate_program_set_const(uint64_t handle, uint32_t index, const uint8_t value[32]) {
    Call call;
    XTRY(call.begin_handle(handle));
    Ext& x = *call.x;
    (void)x;
    auto it = x.programs.find(handle);
    if (it == x.programs.end()) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown gate-program handle");
    if (!value || index >= it->second.n_consts) return ctx::fail(B200ZK_ERR_INVALID_ARG, "gate program: constant index out of range");
    cudaStream_t s = call.stream();
    uint32_t* d = it->second.consts + 8 * (size_t)index;
    XCU(cudaDeviceSynchronize());
    XCU(cudaMemcpyAsync(d, value, 32, cudaMemcpyHostToDevice, s));
    XLAUNCH(fr_convert_kernel2, 1, 32, 0, s, (const uint32_t*)d, d, (uint64_t)1, 1u);
    XCU(cudaStreamSynchronize(s));
    return B200ZK_OK;
}

int32_t b200zk_gate_program_run_dev(uint64_t handle, const void* const* d_columns, void* d_out, uint32_t accumulate, void* stream) {
    Call call;
    XTRY(call.begin_handle(handle));
    Ext& x = *call.x;
    (void)x;
    WsGuard ws;
    XTRY(ws.enter(call.d, S(stream)));
    auto it = x.programs.find(handle);
    if (it == x.programs.end()) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown gate-program handle");
    const GateProgram& gp = it->second;
    if (!d_out || (!d_columns && gp.n_columns)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    if (misaligned(d_out)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "device pointers must be 16-byte aligned");
    for (uint32_t c = 0; c < gp.n_columns; c++)
        if (!d_columns[c] || misaligned(d_columns[c])) return ctx::fail(B200ZK_ERR_INVALID_ARG, "gate program: null or misaligned column " + std::to_string(c));
    cudaStream_t s = S(stream);
    // column pointers ride in a small ring like the host scalars (stream-ordered copy)
    XTRY(x.colptr.ensure(SMALL_BYTES));
    size_t& off = x.colptr_off;
    size_t bytes = ((size_t)gp.n_columns * 8 + 63) & ~(size_t)63;
    if (bytes > SMALL_BYTES / 2) return ctx::fail(B200ZK_ERR_INVALID_ARG, "gate program: too many columns");
    if (off + bytes > SMALL_BYTES) off = 0;
    void* d_ptrs = x.colptr.as<uint8_t>() + off;
    off += bytes;
    if (gp.n_columns) XCU(cudaMemcpyAsync(d_ptrs, d_columns, (size_t)gp.n_columns * 8, cudaMemcpyHostToDevice, s));
    GateArgs a;
    a.columns = reinterpret_cast<const uint32_t* const*>(d_ptrs);
    a.rotations = gp.rotations;
    a.consts = gp.consts;
    a.program = gp.program;
    a.t_inv = gp.t_inv;
    a.out = reinterpret_cast<uint32_t*>(d_out);
    a.n_instr = gp.n_instr;
    a.log_ext = gp.log_ext;
    a.log_period = gp.log_period;
    a.accumulate = accumulate ? 1u : 0u;
    uint64_t n = (uint64_t)1 << gp.log_ext;
    XLAUNCH(fr_gate_eval_kernel, blocks(n, 128), 128, 0, s, a);
    return B200ZK_OK;
}

int32_t b200zk_gate_program_release(uint64_t handle) {
    Call call;
    XTRY(call.begin_handle(handle));
    Ext& x = *call.x;
    (void)x;
    auto it = x.programs.find(handle);
    if (it == x.programs.end()) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown gate-program handle");
    XCU(cudaDeviceSynchronize());
    cudaFree(it->second.program);
    cudaFree(it->second.consts);
    cudaFree(it->second.rotations);
    if (it->second.t_inv) cudaFree(it->second.t_inv);
    x.programs.erase(it);
    return B200ZK_OK;
}

}  // extern "C"
