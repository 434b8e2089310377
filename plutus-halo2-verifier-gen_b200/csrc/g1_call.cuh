// g1_call.cuh -- memory helpers and the out-of-line ("call") forms of the G1 point operations shared by
// the MSM kernels (msm.cuh) and the SRS / decompression kernels (srs.cuh).  Same formulas as g1.cuh
// (replacing blst's G1 add / double / mixed add under midnight_curves::G1Projective,
// /root/reference/Cargo.toml:27); what differs is only how the field products are emitted.
#pragma once
#include "g1.cuh"

namespace b200zk {

// ---------------------------------------------------------------------------------------
// memory helpers (16-byte vector accesses; every G1 array is 16-byte aligned)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ Fp fp_ld(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    Fp r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
__device__ __forceinline__ Fp fp_ldg(const uint32_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    Fp r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
__device__ __forceinline__ void fp_st(uint32_t* p, const Fp& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}
__device__ __forceinline__ G1Affine g1a_ldg(const uint32_t* bases, uint64_t i) {
    G1Affine a;
    a.x = fp_ldg(bases + 24 * i);
    a.y = fp_ldg(bases + 24 * i + 12);
    return a;
}
__device__ __forceinline__ G1Xyzz xyzz_ld(const uint32_t* arr, uint64_t i) {
    G1Xyzz a;
    a.x = fp_ld(arr + 48 * i); a.y = fp_ld(arr + 48 * i + 12);
    a.zz = fp_ld(arr + 48 * i + 24); a.zzz = fp_ld(arr + 48 * i + 36);
    return a;
}
__device__ __forceinline__ void xyzz_st(uint32_t* arr, uint64_t i, const G1Xyzz& a) {
    fp_st(arr + 48 * i, a.x); fp_st(arr + 48 * i + 12, a.y);
    fp_st(arr + 48 * i + 24, a.zz); fp_st(arr + 48 * i + 36, a.zzz);
}

// Out-of-line field product: the point formulas of every kernel except the fully inlined accumulate
// variants call it, so a point addition is ~14 calls plus glue instead of ~75 KB of unrolled
// IMAD chains -- the code stays in the instruction cache and compiles in seconds.
static __device__ __noinline__ Fp fp_mul_call(Fp a, Fp b) { return fe_mul(a, b); }
struct MulCall {
    static __device__ __forceinline__ Fp mul(const Fp& a, const Fp& b) { return fp_mul_call(a, b); }
    static __device__ __forceinline__ Fp sqr(const Fp& a) { return fp_mul_call(a, a); }
    // a*b - c*d
    static __device__ __forceinline__ Fp mulsub(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        return fe_sub(fp_mul_call(a, b), fp_mul_call(c, d));
    }
};
// same, with the two squarings of a mixed addition through the dedicated squaring (fe_sqr_fast: 222 limb
// products instead of 288)
static __device__ __noinline__ Fp fp_sqr_call(Fp a) { return fe_sqr_fast(a); }
static __device__ __noinline__ Fp fp_mul2_call(Fp a, Fp b, Fp c, Fp d) { return fe_mul2(a, b, c, d); }
struct MulCallSqr {
    static __device__ __forceinline__ Fp mul(const Fp& a, const Fp& b) { return fp_mul_call(a, b); }
    static __device__ __forceinline__ Fp sqr(const Fp& a) { return fp_sqr_call(a); }
    // a*b - c*d = a*b + c*(p - d) with one shared Montgomery reduction (fe_mul2: 433 instead of 578 IMAD.WIDE)
    static __device__ __forceinline__ Fp mulsub(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        return fp_mul2_call(a, b, c, fe_neg(d));
    }
};
__device__ __forceinline__ Fp fp_mul_ni(const Fp& a, const Fp& b) { return fp_mul_call(a, b); }
// The tail kernels (collapse, bucket-reduce tree, combine, partial sums) run few threads, so what counts
// there is the latency of ONE point addition.  They all go through the two functions below: operands and
// result in memory (any address space), the 14 (9) field products inlined inside so that independent
// products overlap and no call/spill traffic sits between them.  One copy of that code in the program.
__device__ __forceinline__ G1Xyzz xyzz_ld_gen(const uint32_t* p) {
    G1Xyzz a;
#pragma unroll
    for (int k = 0; k < 12; k++) { a.x.l[k] = p[k]; a.y.l[k] = p[12 + k]; a.zz.l[k] = p[24 + k]; a.zzz.l[k] = p[36 + k]; }
    return a;
}
__device__ __forceinline__ void xyzz_st_gen(uint32_t* p, const G1Xyzz& a) {
#pragma unroll
    for (int k = 0; k < 12; k++) { p[k] = a.x.l[k]; p[12 + k] = a.y.l[k]; p[24 + k] = a.zz.l[k]; p[36 + k] = a.zzz.l[k]; }
}
static __device__ __noinline__ void xyzz_dbl_mem(uint32_t* acc) {
    G1Xyzz a = xyzz_ld_gen(acc);
    xyzz_dbl_t<MulInline>(a);
    xyzz_st_gen(acc, a);
}
static __device__ __noinline__ void xyzz_add_mem(uint32_t* acc, const uint32_t* q) {
    G1Xyzz a = xyzz_ld_gen(acc), b = xyzz_ld_gen(q);
    if (xyzz_is_inf(b)) return;
    if (xyzz_is_inf(a)) { xyzz_st_gen(acc, b); return; }
    Fp u1 = fe_mul(a.x, b.zz), u2 = fe_mul(b.x, a.zz), s1 = fe_mul(a.y, b.zzz), s2 = fe_mul(b.y, a.zzz);
    Fp p = fe_sub(u2, u1), r = fe_sub(s2, s1);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) xyzz_dbl_mem(acc);               // same point (acc is still unchanged in memory)
        else { xyzz_set_inf(a); xyzz_st_gen(acc, a); }       // opposite points
        return;
    }
    Fp pp = fe_sqr_fast(p);
    Fp ppp = fe_mul(p, pp), qq = fe_mul(u1, pp), zz = fe_mul(a.zz, b.zz), zzz = fe_mul(a.zzz, b.zzz);
    Fp x3 = fe_sub(fe_sub(fe_sub(fe_sqr_fast(r), ppp), qq), qq);
    a.y = fe_mul2(r, fe_sub(qq, x3), s1, fe_neg(ppp));   // R (Q - X3) - S1 PPP with one reduction
    a.x = x3;
    a.zz = fe_mul(zz, pp);
    a.zzz = fe_mul(zzz, ppp);
    xyzz_st_gen(acc, a);
}
__device__ __forceinline__ void xyzz_add_ni(G1Xyzz& acc, const G1Xyzz& q) {
    uint32_t ta[48], tq[48];
    xyzz_st_gen(ta, acc);
    xyzz_st_gen(tq, q);
    xyzz_add_mem(ta, tq);
    acc = xyzz_ld_gen(ta);
}
__device__ __forceinline__ void xyzz_dbl_ni(G1Xyzz& acc) {
    uint32_t ta[48];
    xyzz_st_gen(ta, acc);
    xyzz_dbl_mem(ta);
    acc = xyzz_ld_gen(ta);
}
// binary-GCD inversion, out of line (one call site per kernel)
static __device__ __noinline__ Fp fp_inv_ni(Fp a) { return fe_inv_gcd(a); }
__device__ __forceinline__ G1Affine xyzz_to_affine_ni(const G1Xyzz& a) {
    G1Affine r;
    if (xyzz_is_inf(a)) { r.x = fe_zero<FpParams>(); r.y = fe_zero<FpParams>(); return r; }
    Fp t = fp_inv_ni(fp_mul_ni(a.zz, a.zzz));
    r.x = fp_mul_ni(a.x, fp_mul_ni(t, a.zzz));
    r.y = fp_mul_ni(a.y, fp_mul_ni(t, a.zz));
    return r;
}
// acc += (neg ? -q : q); same formulas as xyzz_add_mixed, ordered to keep few values live
template <class M>
__device__ __forceinline__ void xyzz_add_mixed_t(G1Xyzz& acc, const G1Affine& q, bool neg) {
    if (g1a_is_inf(q)) return;
    Fp qy = neg ? fe_neg(q.y) : q.y;
    if (xyzz_is_inf(acc)) {
        acc.x = q.x; acc.y = qy; acc.zz = fe_one<FpParams>(); acc.zzz = fe_one<FpParams>();
        return;
    }
    Fp p = fe_sub(M::mul(q.x, acc.zz), acc.x);
    Fp r = fe_sub(M::mul(qy, acc.zzz), acc.y);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) xyzz_dbl_affine(acc, q.x, qy);  // same point (cold: inlined code that is never fetched)
        else xyzz_set_inf(acc);
        return;
    }
    Fp pp = M::sqr(p);
    Fp qq = M::mul(acc.x, pp);
    acc.zz = M::mul(acc.zz, pp);
    Fp ppp = M::mul(p, pp);
    acc.zzz = M::mul(acc.zzz, ppp);
    Fp x3 = fe_sub(fe_sub(fe_sub(M::sqr(r), ppp), qq), qq);
    acc.y = M::mulsub(r, fe_sub(qq, x3), acc.y, ppp);
    acc.x = x3;
}
// acc += q, both XYZZ, for kernels that are bound by throughput rather than by the latency of one addition (the lowest levels
// of the bucket tree): products out of line like the accumulate kernel's, ordered so that an input dies as early as possible
template <class M>
__device__ __forceinline__ void xyzz_add_tp(G1Xyzz& acc, const G1Xyzz& q) {
    if (xyzz_is_inf(q)) return;
    if (xyzz_is_inf(acc)) { acc = q; return; }
    Fp u1 = M::mul(acc.x, q.zz);
    Fp s1 = M::mul(acc.y, q.zzz);
    Fp p = fe_sub(M::mul(q.x, acc.zz), u1);
    Fp r = fe_sub(M::mul(q.y, acc.zzz), s1);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) xyzz_dbl_t<M>(acc);             // same point
        else xyzz_set_inf(acc);                              // opposite points
        return;
    }
    Fp pp = M::sqr(p);
    acc.zz = M::mul(M::mul(acc.zz, q.zz), pp);
    Fp qq = M::mul(u1, pp);
    Fp ppp = M::mul(p, pp);
    acc.zzz = M::mul(M::mul(acc.zzz, q.zzz), ppp);
    Fp x3 = fe_sub(fe_sub(fe_sub(M::sqr(r), ppp), qq), qq);
    acc.y = M::mulsub(r, fe_sub(qq, x3), s1, ppp);
    acc.x = x3;
}
__device__ __forceinline__ void xyzz_add_mixed_ni(G1Xyzz& acc, const G1Affine& q, bool neg) {
    xyzz_add_mixed_t<MulCall>(acc, q, neg);
}

}  // namespace b200zk
