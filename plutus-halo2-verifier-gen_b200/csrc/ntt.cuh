// ntt.cuh -- Fr number-theoretic transform for sm_100a.
//
// Replaces the butterflies behind midnight_proofs::poly::EvaluationDomain
// {lagrange_to_coeff, coeff_to_extended, extended_to_coeff, coeff_to_lagrange} / best_fft
// (constructed inside keygen/create_proof; explicit call sites
// /root/reference/examples/ivc.rs:109, /root/reference/src/circuits/ivc_circuit.rs:305).
// Convention (pinned by /root/reference/aiken-verifier/aiken_halo2/lib/omega_rotations.ak:48-81):
// X[k] = sum_i x[i] * omega^(i*k), natural order in and out; the caller passes omega.
//
// Structure: a Stockham autosort decomposition into 1..3 passes of radix R = 2^deg (deg <= 11).
// In a pass with current length m and stride s (s*m = n), the R-point sub-transform number
// u = q + s*p reads x[u + j*n/R] (j < R) and writes y[q + s*(R*p + k)] * omega_m^(p*k).
// One CTA stages 2048 elements (= 2048/R sub-transforms with consecutive u, so global reads are
// contiguous across u) in shared memory as 8 word-planes and runs the R-point transform as
// in-place decimation-in-frequency, three radix-2 stages at a time in registers (8 elements
// per thread); the bit reversal of DIF is undone in the address of the final store.
// Per-pass twiddles omega_m^(p*k), the 1/n of an inverse transform and the coset powers are
// table look-ups laid out exactly like the data they multiply.
#pragma once
#include "field.cuh"

namespace b200zk {

constexpr int NTT_LOGB = 11;                  // log2(elements per CTA)
constexpr int NTT_B = 1 << NTT_LOGB;          // 2048 elements = 64 KiB of Fr
constexpr int NTT_THREADS = NTT_B / 8;        // 8 elements per thread
constexpr int NTT_PLANE = NTT_B + NTT_B / 32; // one pad word per 32 elements
constexpr size_t NTT_SMEM = (size_t)8 * NTT_PLANE * sizeof(uint32_t);

struct NttPassArgs {
    const uint32_t* in;       // Fr elements as 8 x u32
    uint32_t* out;
    const uint32_t* tw_local; // omega_R^e, e < R/2 (Montgomery form)
    const uint32_t* tw_pass;  // T[p*R + k] = omega_m^(p*k) (x 1/n where folded) or nullptr
    const uint32_t* in_scale; // per-element multiplier indexed like the input, or nullptr
    const uint32_t* out_scale;// per-element multiplier indexed like the output, or nullptr
    uint32_t log_n;
    uint32_t deg;             // log2 R
    uint32_t log_s;           // log2 of the stride s
    uint32_t log_cols;        // log2(n / R): sub-transforms per polynomial
    uint64_t total_cols;      // batch * n / R
    const uint32_t* scalar;   // multiply every output by *scalar (1/n of a single-pass inverse), or nullptr
    uint32_t reduce_in;       // canonical input may be >= r: reduce on read (transcript.ak:158-179)
    // Multi-GPU four-step transform (b200zk.cu, ntt_sharded): the column pass of GPU g writes element (k1, i2) of the
    // intermediate matrix straight into the HBM of the GPU that owns row k1 (peer store over NVLink) -- the only exchange of
    // the transform is fused into this pass's store; the row pass writes its result transposed ([k2][k1]), which is the
    // layout the strided D2H copy (or the resident caller) wants.
    uint32_t scatter;         // 1: out element (u, k) goes to peer[k >> log_rl] + (((k & (2^log_rl - 1)) << log_c) + col0 + u)
    uint32_t log_rl, log_c, col0;
    uint32_t t_out;           // 0: off; else 1 + log2(batch): out element (poly, oi) goes to (oi << log2 batch) + poly
    uint32_t* peer[16];
};

// Position of element e inside a word plane.  Layout 0: one pad word per 32 elements.  Layout 1 (WL): an XOR swizzle
// that keeps every access pattern of the kernel conflict-free when the lower index bits are owned by single warps:
// bank = e[4:0] ^ (e[7:5] << 2) ^ (e[10:8] << 2) ^ e[7:6]  (lanes may own any five of the bits 0..7, or bits 6..10 in
// the bit-reversed store, and still hit 32 different banks).
template <bool WL> __device__ __forceinline__ uint32_t ntt_pos(uint32_t e) {
    if (WL) return e ^ ((((e >> 5) ^ (e >> 8)) & 7u) << 2) ^ ((e >> 6) & 3u);
    return e + (e >> 5);
}
template <bool WL, int PLANE = NTT_PLANE> __device__ __forceinline__ Fr ntt_lds(const uint32_t* sm, uint32_t e) {
    Fr r;
    uint32_t pos = ntt_pos<WL>(e);
#pragma unroll
    for (int w = 0; w < 8; w++) r.l[w] = sm[w * PLANE + pos];
    return r;
}
template <bool WL, int PLANE = NTT_PLANE> __device__ __forceinline__ void ntt_sts(uint32_t* sm, uint32_t e, const Fr& v) {
    uint32_t pos = ntt_pos<WL>(e);
#pragma unroll
    for (int w = 0; w < 8; w++) sm[w * PLANE + pos] = v.l[w];
}
__device__ __forceinline__ Fr ntt_ldg(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
// data and per-element tables are touched once per pass: streaming loads / stores (evict first), so that the L1 that the shared
// memory carve-out leaves keeps the local twiddle table, which every butterfly reads
__device__ __forceinline__ Fr ntt_ld_stream(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldcs(q), b = __ldcs(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void ntt_st_stream(uint32_t* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    __stcs(q, make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]));
    __stcs(q + 1, make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]));
}
__device__ __forceinline__ Fr ntt_ld(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void ntt_st(uint32_t* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// Fr product, inlined or out of line (smaller loop bodies that stay in the instruction cache)
__device__ __noinline__ Fr fr_mul_call(Fr a, Fr b) { return fe_mul(a, b); }
struct FrMulInline {
    static __device__ __forceinline__ Fr mul(const Fr& a, const Fr& b) { return fe_mul(a, b); }
};
// experiment: every m*r product on the multiplier pipe (FrParamsPlain), same value as fe_mul on Fr
struct FrMulPlain {
    static __device__ __forceinline__ Fr mul(const Fr& a, const Fr& b) {
        Fe<FrParamsPlain> x, y;
#pragma unroll
        for (int i = 0; i < 8; i++) { x.l[i] = a.l[i]; y.l[i] = b.l[i]; }
        Fe<FrParamsPlain> z = fe_mul(x, y);
        Fr r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = z.l[i];
        return r;
    }
};
struct FrMulCall {
    static __device__ __forceinline__ Fr mul(const Fr& a, const Fr& b) { return fr_mul_call(a, b); }
};

// radix-2 DIF stages on index bits lb+NB-1 .. lb of the CTA-local array; the thread owns the 8
// elements whose index differs in bits lb..lb+2.
// TWSM: the R/2 local twiddles were staged in shared memory (word planes like the data) by the caller, so a butterfly never
// waits on a global load.  LB0: this is the sweep over index bits 0..NB-1, where the twiddle exponent of a butterfly depends
// only on its position inside the thread (known at compile time): the multiplications by omega^0 disappear.
template <int NB, class M, bool WL, bool TWSM = false, bool LB0 = false, int LOGB = NTT_LOGB, bool LAZY = false>
__device__ __forceinline__ void ntt_group(uint32_t* sm, const uint32_t* __restrict__ tw_local, uint32_t deg,
                                          uint32_t lb, uint32_t tid) {
    constexpr int PLANE = (1 << LOGB) + (1 << LOGB) / 32;
    if (LB0) lb = 0;
    uint32_t base;
    if (WL && lb + 2 <= 7) {
        // warp-local: warp w owns elements [256 w, 256 w + 256); the lane supplies the five index bits of 0..7 that the
        // thread does not iterate over.  No other warp touches this block until the final store.
        const uint32_t lane = tid & 31, warp = tid >> 5;
        const uint32_t lo = lane & ((1u << lb) - 1), hi = lane >> lb;
        base = (warp << 8) | (hi << (lb + 3)) | lo;
    } else {
        // lanes walk the low index bits when that is conflict-free (lb >= 5), else the high bits
        uint32_t lo, hi;
        if (lb >= 5) {
            lo = tid & ((1u << lb) - 1);
            hi = tid >> lb;
        } else {
            hi = tid & ((1u << (LOGB - 3 - lb)) - 1);
            lo = tid >> (LOGB - 3 - lb);
        }
        base = (hi << (lb + 3)) | lo;
    }
    Fr v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = ntt_lds<WL, PLANE>(sm, base | ((uint32_t)j << lb));
    uint32_t rmask = (1u << deg) - 1;
#pragma unroll
    for (int q = NB - 1; q >= 0; q--) {
        uint32_t b = lb + q;                       // index bit of this stage
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j & (1 << q)) continue;
            Fr x = v[j], y = v[j | (1 << q)];
            v[j] = fe_add(x, y);
            if ((q == 0 && lb == 0) || (LB0 && (j & ((1 << q) - 1)) == 0)) {   // twiddle exponent is 0
                v[j | (1 << q)] = fe_sub(x, y);
            } else {
                // LAZY: the difference goes straight into the product as x - y + r in [0, 2r), without the conditional
                // correction (fe_sub_lazy: 8 LOP3 fewer per butterfly); it is the operand the product walks limb by limb,
                // the twiddle (< r) the full-width one
                const Fr d = LAZY ? fe_sub_lazy(x, y) : fe_sub(x, y);
                uint32_t il = (base | ((uint32_t)j << lb)) & rmask;
                uint32_t e = (il & ((1u << b) - 1)) << (deg - 1 - b);
                if (LB0) e = (uint32_t)(j & ((1 << q) - 1)) << (deg - 1 - b);
                Fr w = TWSM ? ntt_lds<false, PLANE>(sm + 8 * PLANE, e) : ntt_ldg(tw_local + 8 * (size_t)e);
                v[j | (1 << q)] = LAZY ? M::mul(w, d) : M::mul(d, w);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) ntt_sts<WL, PLANE>(sm, base | ((uint32_t)j << lb), v[j]);
}

template <class M, bool WL, bool TWSM = false, bool LB0EN = TWSM, int LOGB = NTT_LOGB, bool LAZY = false>
__device__ __forceinline__ void ntt_pass_body(const NttPassArgs& a) {
    extern __shared__ uint32_t sm[];
    constexpr int PLANE = (1 << LOGB) + (1 << LOGB) / 32, THREADS = (1 << LOGB) / 8;
    const uint32_t tid = threadIdx.x;
    const uint32_t deg = a.deg;
    if (TWSM && deg) {
        // the R/2 local twiddles into planes 8..15 (visible after the barrier that follows the data load)
        for (uint32_t e = tid; e < (1u << (deg - 1)); e += THREADS) ntt_sts<false, PLANE>(sm + 8 * PLANE, e, ntt_ldg(a.tw_local + 8 * (size_t)e));
    }
    const uint32_t logU = LOGB - deg;          // sub-transforms per CTA (log2)
    const uint64_t col0 = (uint64_t)blockIdx.x << logU;
    const uint64_t cmask = ((uint64_t)1 << a.log_cols) - 1;

    // ---- load: element (ul, j) <- x[poly*n + u + j*(n/R)], ul fastest so consecutive lanes read
    //      consecutive u.  The plain path (no scaling) is unrolled so that the eight loads are in flight together;
    //      the scaled paths keep one copy of the product (code size: the kernel must stay in the instruction cache)
    if (!a.in_scale) {
#pragma unroll
        for (int t = 0; t < 8; t++) {
            uint32_t idx = tid + t * THREADS;
            uint32_t ul = idx & ((1u << logU) - 1), j = idx >> logU;
            uint64_t col = col0 + ul;
            Fr v = fe_zero<FrParams>();
            if (col < a.total_cols) {
                uint64_t poly = col >> a.log_cols, u = col & cmask;
                v = ntt_ld_stream(a.in + 8 * ((poly << a.log_n) + u + ((uint64_t)j << a.log_cols)));
                if (a.reduce_in) fe_reduce_loose(v);
            }
            ntt_sts<WL, PLANE>(sm, (ul << deg) | j, v);
        }
    } else
#pragma unroll 1
    for (int t = 0; t < 8; t++) {
        uint32_t idx = tid + t * THREADS;
        uint32_t ul = idx & ((1u << logU) - 1), j = idx >> logU;
        uint64_t col = col0 + ul;
        Fr v = fe_zero<FrParams>();
        if (col < a.total_cols) {
            uint64_t poly = col >> a.log_cols, u = col & cmask;
            uint64_t gi = (poly << a.log_n) + u + ((uint64_t)j << a.log_cols);
            v = ntt_ld_stream(a.in + 8 * gi);
            if (a.reduce_in) fe_reduce_loose(v);
            if (a.in_scale) v = M::mul(v, ntt_ld_stream(a.in_scale + 8 * (u + ((uint64_t)j << a.log_cols))));
        }
        ntt_sts<WL, PLANE>(sm, (ul << deg) | j, v);
    }
    __syncthreads();

    // ---- R-point DIF transforms, three index bits per sweep
    int rem = (int)deg;
    while (rem > 0) {
        int nb = rem >= 3 ? 3 : rem;
        uint32_t lb = (uint32_t)(rem - nb);
        if (LB0EN && lb == 0) {
            if (nb == 3) ntt_group<3, M, WL, TWSM, true, LOGB, LAZY>(sm, a.tw_local, deg, 0, tid);
            else if (nb == 2) ntt_group<2, M, WL, TWSM, true, LOGB, LAZY>(sm, a.tw_local, deg, 0, tid);
            else ntt_group<1, M, WL, TWSM, true, LOGB, LAZY>(sm, a.tw_local, deg, 0, tid);
        } else if (nb == 3) ntt_group<3, M, WL, TWSM, false, LOGB, LAZY>(sm, a.tw_local, deg, lb, tid);
        else if (nb == 2) ntt_group<2, M, WL, TWSM, false, LOGB, LAZY>(sm, a.tw_local, deg, lb, tid);
        else ntt_group<1, M, WL, TWSM, false, LOGB, LAZY>(sm, a.tw_local, deg, lb, tid);
        rem -= nb;
        // the next sweep works on bits below lb: if both this sweep and the next stay inside a warp's 256-element block
        // the warp only has to wait for itself
        if (WL && rem > 0 && lb + 2 <= 7) __syncwarp();
        else __syncthreads();
    }

    // ---- store: y[poly*n + q + s*(R*p + k)] = X_u[k] * T[p*R + k]; X_u[k] sits at bitrev(k).
    //      First pass (s = 1): k fastest.  Later passes: ul (-> q) fastest.
    const bool k_fast = (a.log_s == 0);
#pragma unroll 1
    for (int t = 0; t < 8; t++) {
        uint32_t idx = tid + t * THREADS;
        uint32_t ul, k;
        if (k_fast) { k = idx & ((1u << deg) - 1); ul = idx >> deg; }
        else { ul = idx & ((1u << logU) - 1); k = idx >> logU; }
        uint64_t col = col0 + ul;
        if (col >= a.total_cols) continue;
        uint32_t kr = __brev(k) >> (32 - deg);
        if (deg == 0) kr = 0;
        Fr v = ntt_lds<WL, PLANE>(sm, (ul << deg) | kr);
        uint64_t poly = col >> a.log_cols, u = col & cmask;
        uint64_t q = u & (((uint64_t)1 << a.log_s) - 1), p = u >> a.log_s;
        if (a.tw_pass) v = M::mul(v, ntt_ld_stream(a.tw_pass + 8 * ((p << deg) + k)));
        uint64_t oi = q + (((p << deg) + k) << a.log_s);
        if (a.scatter) {
            ntt_st(a.peer[k >> a.log_rl] + 8 * ((((uint64_t)(k & ((1u << a.log_rl) - 1))) << a.log_c) + a.col0 + u), v);
            continue;
        }
        uint64_t di = a.t_out ? ((oi << (a.t_out - 1)) + poly) : ((poly << a.log_n) + oi);
        if (a.out_scale) v = M::mul(v, ntt_ld_stream(a.out_scale + 8 * (a.t_out ? di : oi)));
        if (a.scalar) v = M::mul(v, ntt_ldg(a.scalar));
        ntt_st_stream(a.out + 8 * di, v);
    }
}

// one CTA per SM (registers unconstrained) and two CTAs per SM (<= 128 registers); the plan picks
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(NttPassArgs a) { ntt_pass_body<FrMulInline, false>(a); }
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_pass_kernel_occ2(NttPassArgs a) { ntt_pass_body<FrMulInline, false>(a); }
// swizzled planes, warp-local sweeps on the index bits 0..7 (CTA barriers only around the sweeps that cross warps)
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_pass_kernel_wl2(NttPassArgs a) { ntt_pass_body<FrMulInline, true>(a); }
// local twiddles staged in shared memory (a second set of word planes), compile-time exponents in the lowest sweep
constexpr size_t NTT_SMEM_TW = (size_t)16 * NTT_PLANE * sizeof(uint32_t);
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_pass_kernel_tw2(NttPassArgs a) { ntt_pass_body<FrMulInline, false, true>(a); }
// compile-time exponents in the lowest sweep only (twiddles stay in global memory / L1)
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_pass_kernel_lb0(NttPassArgs a) { ntt_pass_body<FrMulInline, false, false, true>(a); }
// the same with uncorrected differences in the butterflies (variant 9)
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_pass_kernel_lb0_lazy(NttPassArgs a) { ntt_pass_body<FrMulInline, false, false, true, NTT_LOGB, true>(a); }
// the same with a 4096-element tile (132 KiB of shared memory, 512 threads, one CTA per SM -- the same 16 warps per SM): a
// transform of 2^23 / 2^24 points then takes two passes of radix <= 2^12 instead of three of radix 2^8
constexpr int NTT_LOGB12 = 12;
constexpr int NTT_THREADS12 = (1 << NTT_LOGB12) / 8;
constexpr size_t NTT_SMEM12 = (size_t)8 * ((1 << NTT_LOGB12) + (1 << NTT_LOGB12) / 32) * sizeof(uint32_t);
__global__ void __launch_bounds__(NTT_THREADS12, 1) ntt_pass_kernel_lb0_t12(NttPassArgs a) { ntt_pass_body<FrMulInline, false, false, true, NTT_LOGB12>(a); }
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_pass_kernel_call2(NttPassArgs a) { ntt_pass_body<FrMulCall, false>(a); }
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_pass_kernel_plain2(NttPassArgs a) { ntt_pass_body<FrMulPlain, false>(a); }
__global__ void __launch_bounds__(NTT_THREADS, 3) ntt_pass_kernel_call3(NttPassArgs a) { ntt_pass_body<FrMulCall, false>(a); }

// out[i] = scale * base^(e0 + i*mult mod 2^64), generic table generator (Montgomery form in and out).
// kind 0: exponent = i * mult.  kind 1 (pass table): i = p*R + k, exponent = mult * p * k.
__global__ void fr_powers_kernel(uint32_t* out, const uint32_t* base_p, const uint32_t* scale_p, uint64_t count,
                                 uint64_t mult, uint32_t kind, uint32_t deg) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr base = ntt_ld(base_p);
    uint64_t e;
    if (kind == 0) e = i * mult;
    else {
        uint64_t p = i >> deg, k = i & (((uint64_t)1 << deg) - 1);
        e = mult * p * k;
    }
    Fr acc = scale_p ? ntt_ld(scale_p) : fe_one<FrParams>();
    while (e) {
        if (e & 1) acc = fe_mul(acc, base);
        base = fe_sqr(base);
        e >>= 1;
    }
    ntt_st(out + 8 * i, acc);
}

// out[r * cols + c] = base^(offset + r * row_mul + c * col_mul): coset powers laid out like a GPU's slice of a sharded transform
__global__ void fr_power_grid_kernel(uint32_t* out, const uint32_t* base_p, uint64_t rows, uint64_t cols, uint64_t row_mul,
                                     uint64_t col_mul, uint64_t offset) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    uint64_t e = offset + (i / cols) * row_mul + (i % cols) * col_mul;
    Fr base = ntt_ld(base_p), acc = fe_one<FrParams>();
    while (e) {
        if (e & 1) acc = fe_mul(acc, base);
        base = fe_sqr(base);
        e >>= 1;
    }
    ntt_st(out + 8 * i, acc);
}
// out[u * R + k] = scale * base^((u0 + u) * k): the twiddle block between the two steps of a four-step transform
__global__ void fr_power_block_kernel(uint32_t* out, const uint32_t* base_p, const uint32_t* scale_p, uint64_t count, uint32_t deg,
                                      uint64_t u0) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t e = ((i >> deg) + u0) * (i & (((uint64_t)1 << deg) - 1));
    Fr base = ntt_ld(base_p), acc = scale_p ? ntt_ld(scale_p) : fe_one<FrParams>();
    while (e) {
        if (e & 1) acc = fe_mul(acc, base);
        base = fe_sqr(base);
        e >>= 1;
    }
    ntt_st(out + 8 * i, acc);
}

// out = (Montgomery form of) 1 / 2^log_n ; single thread
__global__ void fr_inv_pow2_kernel(uint32_t* out, uint32_t log_n) {
    Fr two = fe_dbl(fe_one<FrParams>());
    Fr n = fe_pow_u64(two, log_n);
    ntt_st(out, fe_inv(n));
}
// elementwise conversions between canonical and Montgomery form (wire-format helpers)
__global__ void fr_convert_kernel(uint32_t* data, uint64_t count, uint32_t to_mont) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr v = ntt_ld(data + 8 * i);
    if (to_mont) { fe_reduce_loose(v); v = fe_to_mont(v); }
    else v = fe_from_mont(v);
    ntt_st(data + 8 * i, v);
}

}  // namespace b200zk
