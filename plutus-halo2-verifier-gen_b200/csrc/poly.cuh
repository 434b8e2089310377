// poly.cuh -- Fr vector kernels for the polynomial side of the prover that sits between the NTTs and
// the commitments, so that columns can stay resident in HBM from
// coeff_to_extended through the quotient to the h / f / pi commitments:
//
//   * gate-program evaluation over the extended coset domain with the division by X^n - 1 fused in
//     (upstream midnight_proofs::plonk evaluation of h(X); the reference walks the same gate
//     expressions at /root/reference/src/plutus_gen/extraction/mod.rs:81-102 and consumes the result
//     as the "vanishing" commitments, extraction_steps/proof.rs:76-80);
//   * suffix-Horner scan = evaluation at a point + Kate division by (X - z) in one sweep
//     (the q_i / f / pi polynomials of multi_open, /root/reference/src/plutus_gen/extraction/pcs/kzg.rs:55-79,
//     verifier twin /root/reference/aiken-verifier/aiken_halo2/lib/halo2_kzg.ak:46-171);
//   * running products (permutation / lookup grand products, extraction_steps/permutation.rs) and
//     batched inversion (their denominators);
//   * linear combinations and pointwise products of columns.
//
// All data is Fr in Montgomery form (the in-memory form of midnight_curves::Fq), 8 x u32 per element.
// These kernels are bound by HBM or by the integer pipe depending on the products per element; the
// bench prints both fractions.
#pragma once
#include "field.cuh"

namespace b200zk {

__device__ __forceinline__ Fr fr_ld(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr fr_ldg(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void fr_st(uint32_t* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// out-of-line product: the scans and the gate interpreter have many call sites and are not limited by
// the call overhead
__device__ __noinline__ Fr fr_mul_ni(Fr a, Fr b) { return fe_mul(a, b); }
__device__ __noinline__ Fr fr_inv_ni(Fr a) { return fe_inv_fast(a); }

// ---------------------------------------------------------------------------------------
// pointwise operations and format conversion
// ---------------------------------------------------------------------------------------
// op 0: out = a*b   1: out = a+b   2: out = a-b   3: out = a*scalar   4: out = a*b + c (c = out's old value)
__global__ void __launch_bounds__(256) fr_pointwise_kernel(uint32_t op, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                           const uint32_t* __restrict__ scalar, uint32_t* out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = fr_ld(a + 8 * i), r;
    if (op == 0) r = fe_mul(x, fr_ld(b + 8 * i));
    else if (op == 1) r = fe_add(x, fr_ld(b + 8 * i));
    else if (op == 2) r = fe_sub(x, fr_ld(b + 8 * i));
    else if (op == 3) r = fe_mul(x, fr_ldg(scalar));
    else r = fe_add(fe_mul(x, fr_ld(b + 8 * i)), fr_ld(out + 8 * i));
    fr_st(out + 8 * i, r);
}
__global__ void __launch_bounds__(256) fr_convert_kernel2(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t to_mont) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fr_ld(in + 8 * i);
    if (to_mont) { fe_reduce_loose(v); v = fe_to_mont(v); }
    else v = fe_from_mont(v);
    fr_st(out + 8 * i, v);
}

// out[i] = sum_k coeff[k] * polys[k][i], k < count (<= LINCOMB_MAX per launch); accumulate != 0 adds out's old value
constexpr int LINCOMB_MAX = 16;
struct LincombArgs {
    const uint32_t* poly[LINCOMB_MAX];
    uint32_t count;
    uint32_t accumulate;
};
__global__ void __launch_bounds__(256) fr_lincomb_kernel(LincombArgs a, const uint32_t* __restrict__ coeff, uint32_t* out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr acc = a.accumulate ? fr_ld(out + 8 * i) : fe_zero<FrParams>();
    for (uint32_t k = 0; k < a.count; k++) acc = fe_add(acc, fr_mul_ni(fr_ld(a.poly[k] + 8 * i), fr_ldg(coeff + 8 * k)));
    fr_st(out + 8 * i, acc);
}

// out[i] = scale * base^i  (i < n): the powers of the SRS secret, omega^i, coset powers.
// Each thread raises base to its first index once and then walks ITEMS consecutive powers.
__global__ void __launch_bounds__(256) fr_geometric_kernel(const uint32_t* __restrict__ base_p, const uint32_t* __restrict__ scale_p,
                                                           uint32_t* out, uint64_t n) {
    constexpr int ITEMS = 8;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t i0 = t * ITEMS;
    if (i0 >= n) return;
    Fr base = fr_ldg(base_p);
    Fr acc = scale_p ? fr_ldg(scale_p) : fe_one<FrParams>();
    Fr b = base;
    for (uint64_t e = i0; e; e >>= 1) {
        if (e & 1) acc = fr_mul_ni(acc, b);
        b = fr_mul_ni(b, b);
    }
    for (int j = 0; j < ITEMS && i0 + j < n; j++) {
        fr_st(out + 8 * (i0 + j), acc);
        acc = fr_mul_ni(acc, base);
    }
}

// out[r * cols + c] = base^((row0 + r) * c): the twiddle block omega^(i2 * k1) of a four-step / multi-GPU
// transform, one thread per entry (square-and-multiply over the 64-bit exponent)
__global__ void __launch_bounds__(256) fr_power_table_kernel(const uint32_t* __restrict__ base_p, uint64_t row0, uint64_t rows,
                                                             uint64_t cols, uint32_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    uint64_t r = i / cols, c = i - r * cols;
    uint64_t e = (row0 + r) * c;
    Fr b = fr_ldg(base_p), acc = fe_one<FrParams>();
    for (; e; e >>= 1) {
        if (e & 1) acc = fr_mul_ni(acc, b);
        b = fr_mul_ni(b, b);
    }
    fr_st(out + 8 * i, acc);
}

// single-thread helpers: *v <- (*v)^e ;  out <- (s^(2^k) - 1) / 2^k  (the constant of the SRS Lagrange scalars)
__global__ void fr_pow_small_kernel(uint32_t* v, uint32_t e) {
    Fr b = fr_ld(v), acc = fe_one<FrParams>();
    for (; e; e >>= 1) {
        if (e & 1) acc = fr_mul_ni(acc, b);
        b = fr_mul_ni(b, b);
    }
    fr_st(v, acc);
}
__global__ void srs_lagrange_const_kernel(const uint32_t* s_p, uint32_t k, uint32_t* out) {
    Fr s = fr_ld(s_p), n = fe_one<FrParams>();
    for (uint32_t i = 0; i < k; i++) { s = fr_mul_ni(s, s); n = fe_dbl(n); }
    fr_st(out, fr_mul_ni(fe_sub(s, fe_one<FrParams>()), fr_inv_ni(n)));
}

// ---------------------------------------------------------------------------------------
// batched inversion (Montgomery's trick).  One thread owns ITEMS elements strided by the CTA width
// (coalesced), keeps the running products in `scratch` (n elements), inverts one value with the
// branch-free binary-GCD inversion and walks back.  Zeros stay zero, as halo2's batch_invert does.
// ---------------------------------------------------------------------------------------
constexpr int BINV_ITEMS = 32, BINV_THREADS = 128, BINV_TILE = BINV_ITEMS * BINV_THREADS;
__global__ void __launch_bounds__(BINV_THREADS) fr_batch_invert_kernel(const uint32_t* in, uint32_t* out, uint32_t* __restrict__ scratch,
                                                                       uint64_t n) {
    uint64_t base = (uint64_t)blockIdx.x * BINV_TILE + threadIdx.x;
    Fr run = fe_one<FrParams>();
    for (int j = 0; j < BINV_ITEMS; j++) {
        uint64_t i = base + (uint64_t)j * BINV_THREADS;
        if (i >= n) break;
        fr_st(scratch + 8 * i, run);                       // product of the earlier non-zero elements
        Fr v = fr_ld(in + 8 * i);
        if (!fe_is_zero(v)) run = fr_mul_ni(run, v);
    }
    Fr inv = fr_inv_ni(run);                               // run is a product of non-zero values (or one)
    for (int j = BINV_ITEMS - 1; j >= 0; j--) {
        uint64_t i = base + (uint64_t)j * BINV_THREADS;
        if (i >= n) continue;
        Fr v = fr_ld(in + 8 * i);
        if (fe_is_zero(v)) { fr_st(out + 8 * i, v); continue; }
        Fr pre = fr_ld(scratch + 8 * i);
        fr_st(out + 8 * i, fr_mul_ni(inv, pre));
        inv = fr_mul_ni(inv, v);
    }
}

// ---------------------------------------------------------------------------------------
// Scans.  A tile is PS_TILE consecutive elements; a thread owns PS_ITEMS consecutive elements.
//   pass A   : tile totals
//   middle   : exclusive scan of the tile totals by one CTA (in place -> tile carries)
//   pass B   : thread totals -> block scan -> each thread re-walks its elements from its carry
// Two operators:
//   OpMul    : running product, out[i] = init * prod_{j<i} v[j] (exclusive) or prod_{j<=i} (inclusive)
//   OpAffine : composition of the maps x -> z*x + c over the REVERSED coefficient order, i.e. the suffix
//              Horner sums s_i = sum_{j>=i} c_j z^(j-i):  s_0 = p(z), s_{i+1} = coefficient i of the
//              quotient (p(X) - p(z)) / (X - z).
// ---------------------------------------------------------------------------------------
constexpr int PS_ITEMS = 16, PS_THREADS = 128, PS_TILE = PS_ITEMS * PS_THREADS;

struct AffMap {   // x -> m*x + v
    Fr m, v;
};
struct OpMul {
    typedef Fr T;
    static __device__ __forceinline__ T identity() { return fe_one<FrParams>(); }
    static __device__ __forceinline__ T combine(const T& first, const T& second) { return fr_mul_ni(first, second); }
};
struct OpAffine {
    typedef AffMap T;
    static __device__ __forceinline__ T identity() { T t; t.m = fe_one<FrParams>(); t.v = fe_zero<FrParams>(); return t; }
    // apply `first`, then `second`
    static __device__ __forceinline__ T combine(const T& first, const T& second) {
        T t;
        t.m = fr_mul_ni(second.m, first.m);
        t.v = fe_add(fr_mul_ni(second.m, first.v), second.v);
        return t;
    }
};

template <class T> __device__ __forceinline__ T ps_shfl_up(const T& v, int d) {
    T r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 4); k++) o[k] = __shfl_up_sync(0xffffffffu, s[k], d);
    return r;
}
template <class T> __device__ __forceinline__ T ps_shfl_down(const T& v, int d) {
    T r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 4); k++) o[k] = __shfl_down_sync(0xffffffffu, s[k], d);
    return r;
}

// ordered fold of the per-thread totals of a CTA; the result is valid in thread 0
template <class Op> __device__ __forceinline__ typename Op::T ps_block_fold(typename Op::T v, typename Op::T* warp_tot) {
    typedef typename Op::T T;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 1; d < 32; d <<= 1) {
        T o = ps_shfl_down(v, d);
        if (lane + d < 32) v = Op::combine(v, o);
    }
    if (lane == 0) warp_tot[warp] = v;
    __syncthreads();
    T r = Op::identity();
    if (threadIdx.x == 0) {
        r = warp_tot[0];
        for (uint32_t w = 1; w < blockDim.x / 32; w++) r = Op::combine(r, warp_tot[w]);
    }
    return r;
}
// exclusive scan of the per-thread totals of a CTA (identity for thread 0)
template <class Op> __device__ __forceinline__ typename Op::T ps_block_exclusive(typename Op::T v, typename Op::T* warp_tot) {
    typedef typename Op::T T;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 1; d < 32; d <<= 1) {
        T o = ps_shfl_up(v, d);
        if (lane >= d) v = Op::combine(o, v);
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    T ex = ps_shfl_up(v, 1);
    if (lane == 0) ex = Op::identity();
    T pre = Op::identity();
    for (uint32_t w = 0; w < warp; w++) pre = Op::combine(pre, warp_tot[w]);
    return Op::combine(pre, ex);
}

// ---- running product ----
__global__ void __launch_bounds__(PS_THREADS) fr_prodscan_totals_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ totals) {
    __shared__ Fr warp_tot[PS_THREADS / 32];
    uint64_t i0 = (uint64_t)blockIdx.x * PS_TILE + (uint64_t)threadIdx.x * PS_ITEMS;
    Fr run = fe_one<FrParams>();
    for (int j = 0; j < PS_ITEMS; j++)
        if (i0 + j < n) run = fr_mul_ni(run, fr_ld(in + 8 * (i0 + j)));
    Fr tot = ps_block_fold<OpMul>(run, warp_tot);
    if (threadIdx.x == 0) fr_st(totals + 8 * (uint64_t)blockIdx.x, tot);
}
// one CTA: totals[t] <- init * prod_{u<t} totals[u]
__global__ void __launch_bounds__(256) fr_prodscan_middle_kernel(uint32_t* totals, uint64_t count, const uint32_t* __restrict__ init) {
    __shared__ Fr warp_tot[8];
    uint64_t per = (count + blockDim.x - 1) / blockDim.x;
    uint64_t lo = (uint64_t)threadIdx.x * per, hi = lo + per < count ? lo + per : count;
    Fr run = fe_one<FrParams>();
    for (uint64_t t = lo; t < hi; t++) run = fr_mul_ni(run, fr_ld(totals + 8 * t));
    Fr pre = ps_block_exclusive<OpMul>(run, warp_tot);
    if (init) pre = fr_mul_ni(fr_ldg(init), pre);
    for (uint64_t t = lo; t < hi; t++) {
        Fr v = fr_ld(totals + 8 * t);
        fr_st(totals + 8 * t, pre);
        pre = fr_mul_ni(pre, v);
    }
}
__global__ void __launch_bounds__(PS_THREADS) fr_prodscan_apply_kernel(const uint32_t* in, uint32_t* out, uint64_t n,
                                                                       const uint32_t* __restrict__ carries, uint32_t inclusive) {
    __shared__ Fr warp_tot[PS_THREADS / 32];
    uint64_t i0 = (uint64_t)blockIdx.x * PS_TILE + (uint64_t)threadIdx.x * PS_ITEMS;
    Fr run = fe_one<FrParams>();
    for (int j = 0; j < PS_ITEMS; j++)
        if (i0 + j < n) run = fr_mul_ni(run, fr_ld(in + 8 * (i0 + j)));
    Fr pre = ps_block_exclusive<OpMul>(run, warp_tot);
    pre = fr_mul_ni(fr_ld(carries + 8 * (uint64_t)blockIdx.x), pre);
    for (int j = 0; j < PS_ITEMS; j++) {
        if (i0 + j >= n) break;
        Fr v = fr_ld(in + 8 * (i0 + j));                      // read before the (possibly in-place) store
        Fr nxt = fr_mul_ni(pre, v);
        fr_st(out + 8 * (i0 + j), inclusive ? nxt : pre);
        pre = nxt;
    }
}

// ---- suffix Horner (evaluation + Kate division).  Position k of the scan is coefficient n-1-k. ----
// zpow[0] = z, zpow[1] = z^PS_ITEMS (Montgomery form)
__device__ __forceinline__ AffMap horner_thread_total(const uint32_t* __restrict__ coeffs, uint64_t n, uint64_t k0, const Fr& z, const Fr& zI) {
    Fr v = fe_zero<FrParams>();
    for (int j = 0; j < PS_ITEMS; j++) {
        uint64_t k = k0 + j;
        Fr c = k < n ? fr_ld(coeffs + 8 * (n - 1 - k)) : fe_zero<FrParams>();   // padding below X^0 is never stored
        v = fe_add(fr_mul_ni(v, z), c);
    }
    AffMap t;
    t.m = zI;
    t.v = v;
    return t;
}
__global__ void __launch_bounds__(PS_THREADS) fr_horner_totals_kernel(const uint32_t* __restrict__ coeffs, uint64_t n,
                                                                      const uint32_t* __restrict__ zpow, uint32_t* __restrict__ totals) {
    __shared__ AffMap warp_tot[PS_THREADS / 32];
    uint64_t k0 = (uint64_t)blockIdx.x * PS_TILE + (uint64_t)threadIdx.x * PS_ITEMS;
    AffMap t = horner_thread_total(coeffs, n, k0, fr_ldg(zpow), fr_ldg(zpow + 8));
    AffMap tot = ps_block_fold<OpAffine>(t, warp_tot);
    if (threadIdx.x == 0) { fr_st(totals + 16 * (uint64_t)blockIdx.x, tot.m); fr_st(totals + 16 * (uint64_t)blockIdx.x + 8, tot.v); }
}
// one CTA: totals[t] <- composition of the maps of tiles u < t (applied to 0: only .v is needed later, .m kept too)
__global__ void __launch_bounds__(256) fr_horner_middle_kernel(uint32_t* totals, uint64_t count) {
    __shared__ AffMap warp_tot[8];
    uint64_t per = (count + blockDim.x - 1) / blockDim.x;
    uint64_t lo = (uint64_t)threadIdx.x * per, hi = lo + per < count ? lo + per : count;
    AffMap run = OpAffine::identity();
    for (uint64_t t = lo; t < hi; t++) {
        AffMap e;
        e.m = fr_ld(totals + 16 * t);
        e.v = fr_ld(totals + 16 * t + 8);
        run = OpAffine::combine(run, e);
    }
    AffMap pre = ps_block_exclusive<OpAffine>(run, warp_tot);
    for (uint64_t t = lo; t < hi; t++) {
        AffMap e;
        e.m = fr_ld(totals + 16 * t);
        e.v = fr_ld(totals + 16 * t + 8);
        fr_st(totals + 16 * t, pre.m);
        fr_st(totals + 16 * t + 8, pre.v);
        pre = OpAffine::combine(pre, e);
    }
}
// quot (n-1 coefficients, may be null) and eval (one element, may be null)
__global__ void __launch_bounds__(PS_THREADS) fr_horner_apply_kernel(const uint32_t* __restrict__ coeffs, uint64_t n,
                                                                     const uint32_t* __restrict__ zpow, const uint32_t* __restrict__ carries,
                                                                     uint32_t* __restrict__ quot, uint32_t* __restrict__ eval) {
    __shared__ AffMap warp_tot[PS_THREADS / 32];
    uint64_t k0 = (uint64_t)blockIdx.x * PS_TILE + (uint64_t)threadIdx.x * PS_ITEMS;
    Fr z = fr_ldg(zpow);
    AffMap t = horner_thread_total(coeffs, n, k0, z, fr_ldg(zpow + 8));
    AffMap pre = ps_block_exclusive<OpAffine>(t, warp_tot);
    // value entering this thread = (tile carry, then the earlier threads of the tile) applied to 0
    Fr s = fe_add(fr_mul_ni(pre.m, fr_ld(carries + 16 * (uint64_t)blockIdx.x + 8)), pre.v);
    for (int j = 0; j < PS_ITEMS; j++) {
        uint64_t k = k0 + j;
        if (k >= n) break;
        uint64_t i = n - 1 - k;
        s = fe_add(fr_mul_ni(s, z), fr_ld(coeffs + 8 * i));   // s_i
        if (i == 0) { if (eval) fr_st(eval, s); }
        else if (quot) fr_st(quot + 8 * (i - 1), s);
    }
}

// ---------------------------------------------------------------------------------------
// Gate-program evaluation over the extended domain.
//
// A program is a list of register-machine instructions, the shape of halo2's GraphEvaluator
// (calculations over value sources).  One thread evaluates the whole program for one row of the
// extended domain; columns are read at (row + rotation * 2^(ext_k - k)) mod 2^ext_k.  The value of the
// last instruction's destination register, times t_inv[row mod period] (1 / (X^n - 1) on the coset,
// which takes only 2^(ext_k - k) distinct values), is the quotient's evaluation at that row.
//
// instruction = 4 x u32: { op | dst << 8, src a, src b, src c }
//   src = kind << 28 | payload:  kind 0 constant[payload], 1 register[payload],
//                                 kind 2 column[payload >> 12] at rotation index (payload & 0xfff)
//   op  0 add (a+b)  1 sub (a-b)  2 mul (a*b)  3 neg (-a)  4 double (2a)  5 square (a*a)
//       6 muladd (a*b + c: a Horner step)  7 mov (a)
// ---------------------------------------------------------------------------------------
constexpr int GATE_MAX_REGS = 48;
struct GateArgs {
    const uint32_t* const* columns;   // device array of column pointers (Montgomery Fr, 2^log_ext elements each)
    const int32_t* rotations;         // device array: row offsets, already scaled to the extended domain
    const uint32_t* consts;           // device array of Montgomery constants (challenges included)
    const uint32_t* program;          // device array, 4 words per instruction
    const uint32_t* t_inv;            // device array of 2^log_period Montgomery elements, or null
    uint32_t* out;
    uint32_t n_instr;
    uint32_t log_ext;
    uint32_t log_period;
    uint32_t accumulate;              // out[row] += value instead of out[row] = value
};
__device__ __forceinline__ Fr gate_src(uint32_t src, const GateArgs& a, const Fr* regs, uint64_t row, uint64_t mask) {
    uint32_t kind = src >> 28, pay = src & 0x0fffffffu;
    if (kind == 1) return regs[pay];
    if (kind == 0) return fr_ldg(a.consts + 8 * (size_t)pay);
    const uint32_t* col = a.columns[pay >> 12];
    int64_t r = (int64_t)row + (int64_t)a.rotations[pay & 0xfffu];
    return fr_ld(col + 8 * ((uint64_t)r & mask));
}
__global__ void __launch_bounds__(128) fr_gate_eval_kernel(GateArgs a) {
    const uint64_t n = (uint64_t)1 << a.log_ext, mask = n - 1;
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    Fr regs[GATE_MAX_REGS];
    uint32_t last = 0;
    for (uint32_t pc = 0; pc < a.n_instr; pc++) {
        uint4 ins = __ldg(reinterpret_cast<const uint4*>(a.program) + pc);
        uint32_t op = ins.x & 0xffu, dst = ins.x >> 8;
        Fr x = gate_src(ins.y, a, regs, row, mask), r;
        switch (op) {
            case 0: r = fe_add(x, gate_src(ins.z, a, regs, row, mask)); break;
            case 1: r = fe_sub(x, gate_src(ins.z, a, regs, row, mask)); break;
            case 2: r = fr_mul_ni(x, gate_src(ins.z, a, regs, row, mask)); break;
            case 3: r = fe_neg(x); break;
            case 4: r = fe_dbl(x); break;
            case 5: r = fr_mul_ni(x, x); break;
            case 6: r = fe_add(fr_mul_ni(x, gate_src(ins.z, a, regs, row, mask)), gate_src(ins.w, a, regs, row, mask)); break;
            default: r = x; break;
        }
        regs[dst] = r;
        last = dst;
    }
    Fr v = a.n_instr ? regs[last] : fe_zero<FrParams>();
    if (a.t_inv) v = fr_mul_ni(v, fr_ldg(a.t_inv + 8 * (row & (((uint64_t)1 << a.log_period) - 1))));
    if (a.accumulate) v = fe_add(v, fr_ld(a.out + 8 * row));
    fr_st(a.out + 8 * row, v);
}

}  // namespace b200zk
