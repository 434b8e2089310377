// field.cuh -- Montgomery arithmetic for the two BLS12-381 fields on 32-bit limbs,
// written for the sm_100a integer pipe (IMAD.WIDE with carry).
//
//   Fp : 381-bit base field,   12 x u32, R = 2^384  (replaces blst fp, reached through
//        midnight_curves::Fp -- /root/reference/Cargo.toml:27, used at
//        /root/reference/src/plutus_gen/proof_serialization.rs:43,58-59)
//   Fr : 255-bit scalar field,  8 x u32, R = 2^256  (midnight_curves::Fq, the type of every
//        polynomial coefficient -- /root/reference/examples/simple_mul.rs:7)
//
// The limb layout of a Montgomery-form element is byte-identical to blst's 6x/4x u64
// little-endian limbs, so the "MONT" wire formats of include/b200zk.h are zero-copy.
//
// Multiplication uses the even/odd column split: the products a[j]*b_i for even j land on
// 64-bit pairs (0,1),(2,3).. of one accumulator and for odd j on the pairs of a second
// accumulator that is aligned one limb higher, so every row is a run of IMAD.WIDE.U32
// with the carry riding on a predicate, and the division by 2^32 of each Montgomery step
// is a swap of the two accumulators instead of a register shuffle.
//
// Every function is __host__ __device__: on the host the carry flag is emulated, which
// lets tests/host_kernels_test.cu run the very same algorithm against the oracle on a box without
// a GPU.  The product never runs the host path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace b200zk {

// ---------------------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define B200ZK_CC_DECL
HD uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
HD uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// (hi:lo) = a*b + (chi:clo), carry-out to CC.  ptxas fuses each lo/hi pair into one IMAD.WIDE.U32.
HD void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// same with carry-in from CC
HD void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
HD void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    // one IMAD.WIDE.U32 (a separate mul.lo / mul.hi pair is not fused by ptxas and IMAD.HI costs as much as the wide form)
    uint64_t r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    lo = (uint32_t)r;
    hi = (uint32_t)(r >> 32);
}
// column accumulation (c2:c1:c0) += a*b, three instructions: IMAD (carry out), IMAD.HI.X, IADD3.X
HD void mad_col(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                 : "+r"(c0), "+r"(c1), "+r"(c2) : "r"(a), "r"(b));
}
#else
// host emulation of the PTX carry flag (tests only)
static thread_local uint32_t t_cc = 0;
HD uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; t_cc = (uint32_t)(s >> 32); return (uint32_t)s; }
HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + t_cc; t_cc = (uint32_t)(s >> 32); return (uint32_t)s; }
HD uint32_t addc(uint32_t a, uint32_t b) { return a + b + t_cc; }
HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; t_cc = (uint32_t)(s >> 32) & 1; return (uint32_t)s; }
HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - t_cc; t_cc = (uint32_t)(s >> 32) & 1; return (uint32_t)s; }
HD uint32_t subc(uint32_t a, uint32_t b) { return a - b - t_cc; }
HD void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    uint64_t pr = (uint64_t)a * b;
    uint64_t l = (pr & 0xffffffffu) + clo;
    uint64_t h = (pr >> 32) + chi + (l >> 32);
    lo = (uint32_t)l; hi = (uint32_t)h; t_cc = (uint32_t)(h >> 32);
}
HD void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    uint64_t pr = (uint64_t)a * b;
    uint64_t l = (pr & 0xffffffffu) + clo + t_cc;
    uint64_t h = (pr >> 32) + chi + (l >> 32);
    lo = (uint32_t)l; hi = (uint32_t)h; t_cc = (uint32_t)(h >> 32);
}
HD void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    uint64_t pr = (uint64_t)a * b; lo = (uint32_t)pr; hi = (uint32_t)(pr >> 32);
}
HD void mad_col(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
    uint64_t pr = (uint64_t)a * b;
    uint64_t l = (uint64_t)c0 + (uint32_t)pr;
    uint64_t h = (uint64_t)c1 + (pr >> 32) + (l >> 32);
    c0 = (uint32_t)l; c1 = (uint32_t)h; c2 += (uint32_t)(h >> 32);
}
#endif

// ---------------------------------------------------------------------------------------
// field parameters
// ---------------------------------------------------------------------------------------
#define B200ZK_FP_P   {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, \
                       0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
#define B200ZK_FP_PM2 {0xffffaaa9u, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, \
                       0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
#define B200ZK_FP_R   {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u, \
                       0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u}
#define B200ZK_FP_R2  {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu, \
                       0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u}
#define B200ZK_FR_P   {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u}
#define B200ZK_FR_PM2 {0xffffffffu, 0xfffffffeu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u}
#define B200ZK_FR_R   {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u}
#define B200ZK_FR_R2  {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u}

// BLS12-381 base-field modulus p (BlsTypes.hs:102-103 / bls_utils.ak:14-15) and scalar-field
// modulus r (BlsTypes.hs:97), with R mod m, R^2 mod m and m-2.  Device code reads the
// __constant__ copies (they fold into IMAD operands as c[bank][off]); the host copies exist
// only for the CPU-side algorithm tests.
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t c_fp_p[12] = B200ZK_FP_P, c_fp_pm2[12] = B200ZK_FP_PM2, c_fp_r[12] = B200ZK_FP_R, c_fp_r2[12] = B200ZK_FP_R2;
static __device__ __constant__ uint32_t c_fr_p[8] = B200ZK_FR_P, c_fr_pm2[8] = B200ZK_FR_PM2, c_fr_r[8] = B200ZK_FR_R, c_fr_r2[8] = B200ZK_FR_R2;
#endif
static const uint32_t h_fp_p[12] = B200ZK_FP_P, h_fp_pm2[12] = B200ZK_FP_PM2, h_fp_r[12] = B200ZK_FP_R, h_fp_r2[12] = B200ZK_FP_R2;
static const uint32_t h_fr_p[8] = B200ZK_FR_P, h_fr_pm2[8] = B200ZK_FR_PM2, h_fr_r[8] = B200ZK_FR_R, h_fr_r2[8] = B200ZK_FR_R2;

#ifdef __CUDA_ARCH__
#define B200ZK_SEL(name) c_##name
#else
#define B200ZK_SEL(name) h_##name
#endif

struct FpParams {
    static constexpr int N = 12;
    static constexpr uint32_t M0 = 0xfffcfffdu;  // -p^-1 mod 2^32
    static constexpr bool LOW_LIMBS_TRIVIAL = false;
    HD static uint32_t p(int i) { return B200ZK_SEL(fp_p)[i]; }
    HD static uint32_t pm2(int i) { return B200ZK_SEL(fp_pm2)[i]; }
    HD static uint32_t r(int i) { return B200ZK_SEL(fp_r)[i]; }
    HD static uint32_t r2(int i) { return B200ZK_SEL(fp_r2)[i]; }
};
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t M0 = 0xffffffffu;
    static constexpr bool LOW_LIMBS_TRIVIAL = true;   // r = ...ffffffff00000001
    // the modulus limbs are literals: inside the unrolled Montgomery rows they become immediates
    HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[8] = B200ZK_FR_P;
        return v[i];
    }
    HD static uint32_t pm2(int i) { return B200ZK_SEL(fr_pm2)[i]; }
    HD static uint32_t r(int i) { return B200ZK_SEL(fr_r)[i]; }
    HD static uint32_t r2(int i) { return B200ZK_SEL(fr_r2)[i]; }
};

// ---------------------------------------------------------------------------------------
// Field element: fully reduced, Montgomery form
// ---------------------------------------------------------------------------------------
template <class P>
struct Fe {
    static constexpr int N = P::N;
    uint32_t l[N];
};

template <class P> HD Fe<P> fe_zero() { Fe<P> r; for (int i = 0; i < P::N; i++) r.l[i] = 0; return r; }
template <class P> HD Fe<P> fe_one() { Fe<P> r; for (int i = 0; i < P::N; i++) r.l[i] = P::r(i); return r; }
template <class P> HD bool fe_is_zero(const Fe<P>& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < P::N; i++) o |= a.l[i];
    return o == 0;
}
template <class P> HD bool fe_eq(const Fe<P>& a, const Fe<P>& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < P::N; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}

// r = a - p if a >= p else a   (a < 2p)
template <class P> HD void fe_cond_sub_p(Fe<P>& a) {
    constexpr int N = P::N;
    uint32_t t[N];
    t[0] = sub_cc(a.l[0], P::p(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = subc_cc(a.l[i], P::p(i));
    uint32_t borrow = subc(0, 0);  // 0xffffffff if a < p
#pragma unroll
    for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : t[i];
}

template <class P> HD Fe<P> fe_add(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    Fe<P> r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    // both moduli leave the top bit of the top limb clear, so a+b < 2p never carries out
    fe_cond_sub_p(r);
    return r;
}
template <class P> HD Fe<P> fe_sub(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    Fe<P> r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
    uint32_t mask = subc(0, 0);  // all ones on borrow
    r.l[0] = add_cc(r.l[0], P::p(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], P::p(i) & mask);
    r.l[N - 1] = addc(r.l[N - 1], P::p(N - 1) & mask);
    return r;
}
// a - b + p in [0, 2p) for reduced a, b, without the conditional correction (two plain carry chains, no mask): for a difference
// that goes straight into a product as the operand fe_mul walks limb by limb (its SECOND argument).  With the other operand
// below p the running value of the interleaved reduction stays below (other + p) < 2p and the result below p (1 + 2p/R) < 2p, so
// fe_mul's single conditional subtraction still returns the canonical residue.  Needs 2p < 2^(32 N): true for both fields.
template <class P> HD Fe<P> fe_sub_lazy(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    Fe<P> t, r;
    t.l[0] = sub_cc(P::p(0), b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) t.l[i] = subc_cc(P::p(i), b.l[i]);
    t.l[N - 1] = subc(P::p(N - 1), b.l[N - 1]);
    r.l[0] = add_cc(a.l[0], t.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(a.l[i], t.l[i]);
    r.l[N - 1] = addc(a.l[N - 1], t.l[N - 1]);
    return r;
}
template <class P> HD Fe<P> fe_neg(const Fe<P>& a) { return fe_sub(fe_zero<P>(), a); }
template <class P> HD Fe<P> fe_dbl(const Fe<P>& a) { return fe_add(a, a); }

// E += m * p(even limbs), O += m * p(odd limbs) with m chosen so that E[0] becomes 0; the carry that leaves E
// lands on O[N-1] (O is aligned one limb higher).  For Fr the two low limbs of the modulus are 0x00000001 and
// 0xffffffff: their products are additions (m*1 = m; m*(2^32-1) = (m - [m != 0]) * 2^32 + (2^32 - m) mod 2^32),
// which takes 2 of the 8 products of every row off the multiplier pipe.
template <class P> HD void mont_mp_rows(uint32_t* E, uint32_t* O) {
    constexpr int N = P::N;
    uint32_t m;
    if (P::LOW_LIMBS_TRIVIAL) {
        // For M0 = 2^32 - 1 the quotient digit is m = -E[0].  (Written as E[0] * M0 the compiler folds the negation into the
        // products that follow and ptxas then emits every carry-chained m*p product as IMAD.X + IMAD.HI.U32.X, 5 pipe
        // cycles, instead of one IMAD.WIDE.U32.X, 4; the opaque subtraction keeps the rows fused.)
        // The product m * 0xffffffff = (m - c0) * 2^32 + (2^32 - m) with c0 = [m != 0]: the subtraction that forms m leaves
        // exactly that borrow behind, so `hi` is one subtract-with-borrow, no compare / select.  A borrow may only feed a
        // subtract and a carry only an add: PTX's flag after sub.cc is "not borrow" on the device (a + ~b + 1), so the two
        // must not be mixed -- hence E[0] + m is formed by its own add, whose carry (again c0) goes on into the O row.
        const uint32_t e0 = E[0];
        m = sub_cc(0, e0);
        const uint32_t hi = subc(m, 0);        // m - c0
        E[0] = add_cc(e0, m);                  // = 0, carry c0: a unit of limb 1 -- which is O[0]'s weight too (O is aligned one
        O[0] = addc_cc(O[0], e0);              // limb higher), so it enters the O row directly; 2^32 - m = E[0] (mod 2^32)
        O[1] = addc_cc(O[1], hi);
#pragma unroll
        for (int k = 2; k < N; k += 2) madc_wide_cc(O[k], O[k + 1], P::p(k + 1), m, O[k], O[k + 1]);
        // E row: limbs 0 and 1 are done (E[0] = 0, E[1] unchanged); the carry that leaves E lands on O[N-1]
        mad_wide_cc(E[2], E[3], P::p(2), m, E[2], E[3]);
#pragma unroll
        for (int j = 4; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], P::p(j), m, E[j], E[j + 1]);
        O[N - 1] = addc(O[N - 1], 0);
        return;
    }
    m = E[0] * P::M0;
    mad_wide_cc(O[0], O[1], P::p(1), m, O[0], O[1]);
#pragma unroll
    for (int k = 2; k < N; k += 2) madc_wide_cc(O[k], O[k + 1], P::p(k + 1), m, O[k], O[k + 1]);
    mad_wide_cc(E[0], E[1], P::p(0), m, E[0], E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], P::p(j), m, E[j], E[j + 1]);
    O[N - 1] = addc(O[N - 1], 0);
}

// One Montgomery step on the (E, O) accumulator pair: E += a_even*bi, O' = (old E >> 64) + a_odd*bi,
// then add m*p so that E[0] becomes 0.  On entry E is the accumulator that was "odd" in the
// previous step and O the one that was "even" (its limb 0 is already zero, limb 1 still pending).
template <class P> HD void mont_step(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
    E[0] = add_cc(E[0], O[1]);
#pragma unroll
    for (int k = 0; k < N - 2; k += 2) madc_wide_cc(O[k], O[k + 1], a[k + 1], bi, O[k + 2], O[k + 3]);
    madc_wide_cc(O[N - 2], O[N - 1], a[N - 1], bi, 0, 0);
    mad_wide_cc(E[0], E[1], a[0], bi, E[0], E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], a[j], bi, E[j], E[j + 1]);
    O[N - 1] = addc(O[N - 1], 0);
    mont_mp_rows<P>(E, O);
}

// r = a*b/R mod p, inputs and output fully reduced
template <class P> HD Fe<P> fe_mul(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    uint32_t ev[N], od[N];
    // step 0: plain products, then the m*p row
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        mul_wide(ev[j], ev[j + 1], a.l[j], b.l[0]);
        mul_wide(od[j], od[j + 1], a.l[j + 1], b.l[0]);
    }
    mont_mp_rows<P>(ev, od);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_step<P>(od, ev, a.l, b.l[i]);
        if (i + 1 < N) mont_step<P>(ev, od, a.l, b.l[i + 1]);
    }
    // N is even: the last step ran with E = od, O = ev, so the value is ev + (od >> 32)
    Fe<P> r;
    r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(ev[i], od[i + 1]);
    r.l[N - 1] = addc(ev[N - 1], 0);
    fe_cond_sub_p(r);
    return r;
}
// r = (a*b + c*d)/R mod p in one interleaved pass: both product rows are accumulated before the shared m*p row, so
// the pair costs 3 rows per step instead of 4 (433 instead of 578 IMAD.WIDE for Fp).  Fp only: the running value stays
// below 3p(1 + 2^-32), which needs p < 2^(32N) / 3 -- true for the 381-bit p in 384 bits, false for the 255-bit r in
// 256 bits (the host test shows the overflow) -- and the result is below p(1 + 2p/R) < 2p, so one conditional
// subtraction suffices.
template <class P> HD void mont_step2(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, const uint32_t* c, uint32_t di) {
    constexpr int N = P::N;
    E[0] = add_cc(E[0], O[1]);
#pragma unroll
    for (int k = 0; k < N - 2; k += 2) madc_wide_cc(O[k], O[k + 1], a[k + 1], bi, O[k + 2], O[k + 3]);
    madc_wide_cc(O[N - 2], O[N - 1], a[N - 1], bi, 0, 0);
    mad_wide_cc(E[0], E[1], a[0], bi, E[0], E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], a[j], bi, E[j], E[j + 1]);
    O[N - 1] = addc(O[N - 1], 0);
    // second product row on the same (already shifted) accumulators
    mad_wide_cc(O[0], O[1], c[1], di, O[0], O[1]);
#pragma unroll
    for (int k = 2; k < N; k += 2) madc_wide_cc(O[k], O[k + 1], c[k + 1], di, O[k], O[k + 1]);
    mad_wide_cc(E[0], E[1], c[0], di, E[0], E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], c[j], di, E[j], E[j + 1]);
    O[N - 1] = addc(O[N - 1], 0);
    mont_mp_rows<P>(E, O);
}
template <class P> HD Fe<P> fe_mul2(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
    constexpr int N = P::N;
    static_assert(P::N == 12, "fe_mul2 needs three bits of headroom above the modulus (Fp only)");
    uint32_t ev[N], od[N];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        mul_wide(ev[j], ev[j + 1], a.l[j], b.l[0]);
        mul_wide(od[j], od[j + 1], a.l[j + 1], b.l[0]);
    }
    mad_wide_cc(od[0], od[1], c.l[1], d.l[0], od[0], od[1]);
#pragma unroll
    for (int k = 2; k < N; k += 2) madc_wide_cc(od[k], od[k + 1], c.l[k + 1], d.l[0], od[k], od[k + 1]);
    mad_wide_cc(ev[0], ev[1], c.l[0], d.l[0], ev[0], ev[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(ev[j], ev[j + 1], c.l[j], d.l[0], ev[j], ev[j + 1]);
    od[N - 1] = addc(od[N - 1], 0);
    mont_mp_rows<P>(ev, od);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_step2<P>(od, ev, a.l, b.l[i], c.l, d.l[i]);
        if (i + 1 < N) mont_step2<P>(ev, od, a.l, b.l[i + 1], c.l, d.l[i + 1]);
    }
    Fe<P> r;
    r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(ev[i], od[i + 1]);
    r.l[N - 1] = addc(ev[N - 1], 0);
    fe_cond_sub_p(r);
    return r;
}
template <class P> HD Fe<P> fe_sqr(const Fe<P>& a) { return fe_mul(a, a); }

// Dedicated Montgomery squaring: the N(N-1)/2 off-diagonal products are formed once, doubled, the N diagonal squares added, and the 2N-limb square is
// reduced by N rows of IMAD.WIDE chains whose end-of-row carries are deferred (they all land above limb
// N, so the quotient digits m do not depend on them).  78 + 144 limb products instead of 288: the
// multiplier pipe is busy 23 % less than for fe_mul(a, a).  Same fully reduced result.
template <class P> HD Fe<P> fe_sqr_fast(const Fe<P>& a) {
    constexpr int N = P::N;
    // off-diagonal products a_i*a_j (i < j) as IMAD.WIDE carry chains: a product at limb position i+j lands
    // on the pair-aligned accumulator of that parity (A0: pairs at even positions, A1: at odd positions), so
    // the products of one row with the same parity of j sit on adjacent pairs; the carry that leaves a chain
    // is parked in cx and folded in once at the end.
    uint32_t A0[2 * N], A1[2 * N], cx[2 * N + 2];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) { A0[k] = 0; A1[k] = 0; }
#pragma unroll
    for (int k = 0; k < 2 * N + 2; k++) cx[k] = 0;
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
        // j = i+1, i+3, ...: positions 2i+1, 2i+3, ... (odd) -> A1
        mad_wide_cc(A1[2 * i + 1], A1[2 * i + 2], a.l[i], a.l[i + 1], A1[2 * i + 1], A1[2 * i + 2]);
        int last = 2 * i + 1;
#pragma unroll
        for (int j = i + 3; j < N; j += 2) {
            madc_wide_cc(A1[i + j], A1[i + j + 1], a.l[i], a.l[j], A1[i + j], A1[i + j + 1]);
            last = i + j;
        }
        cx[last + 2] = addc(cx[last + 2], 0);
        // j = i+2, i+4, ...: positions 2i+2, 2i+4, ... (even) -> A0
        if (i + 2 < N) {
            mad_wide_cc(A0[2 * i + 2], A0[2 * i + 3], a.l[i], a.l[i + 2], A0[2 * i + 2], A0[2 * i + 3]);
            last = 2 * i + 2;
#pragma unroll
            for (int j = i + 4; j < N; j += 2) {
                madc_wide_cc(A0[i + j], A0[i + j + 1], a.l[i], a.l[j], A0[i + j], A0[i + j + 1]);
                last = i + j;
            }
            cx[last + 2] = addc(cx[last + 2], 0);
        }
    }
    // U = A0 + A1 + cx, T = 2U + sum a_i^2 2^(64 i)
    uint32_t t[2 * N];
    t[0] = add_cc(A0[0], A1[0]);
#pragma unroll
    for (int k = 1; k < 2 * N; k++) t[k] = addc_cc(A0[k], A1[k]);
    t[0] = add_cc(t[0], cx[0]);
#pragma unroll
    for (int k = 1; k < 2 * N; k++) t[k] = addc_cc(t[k], cx[k]);
#pragma unroll
    for (int k = 2 * N - 1; k > 0; k--) t[k] = (t[k] << 1) | (t[k - 1] >> 31);
    t[0] <<= 1;
    {
        uint32_t d0, d1;
        mul_wide(d0, d1, a.l[0], a.l[0]);
        t[0] = add_cc(t[0], d0);
        t[1] = addc_cc(t[1], d1);
#pragma unroll
        for (int i = 1; i < N; i++) {
            mul_wide(d0, d1, a.l[i], a.l[i]);
            t[2 * i] = addc_cc(t[2 * i], d0);
            t[2 * i + 1] = addc_cc(t[2 * i + 1], d1);
        }
    }
    uint32_t ex[N + 1];
#pragma unroll
    for (int i = 0; i <= N; i++) ex[i] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t m = t[i] * P::M0;
        mad_wide_cc(t[i], t[i + 1], P::p(0), m, t[i], t[i + 1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(t[i + j], t[i + j + 1], P::p(j), m, t[i + j], t[i + j + 1]);
        ex[i] = addc(ex[i], 0);                 // carry out of limbs (i+N-2, i+N-1) lands on limb i+N
        mad_wide_cc(t[i + 1], t[i + 2], P::p(1), m, t[i + 1], t[i + 2]);
#pragma unroll
        for (int j = 3; j < N; j += 2) {
            if (i + j + 1 < 2 * N) madc_wide_cc(t[i + j], t[i + j + 1], P::p(j), m, t[i + j], t[i + j + 1]);
        }
        ex[i + 1] = addc(ex[i + 1], 0);         // carry out of limbs (i+N-1, i+N) lands on limb i+N+1
    }
    Fe<P> r;
    r.l[0] = add_cc(t[N], ex[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = addc_cc(t[N + i], ex[i]);
    fe_cond_sub_p(r);
    return r;
}

// canonical (non-Montgomery) limbs -> Montgomery form, and back
template <class P> HD Fe<P> fe_to_mont(const Fe<P>& a) {
    Fe<P> r2;
    for (int i = 0; i < P::N; i++) r2.l[i] = P::r2(i);
    return fe_mul(a, r2);
}
template <class P> HD Fe<P> fe_from_mont(const Fe<P>& a) {
    Fe<P> one = fe_zero<P>();
    one.l[0] = 1;
    return fe_mul(a, one);
}
// reduce an arbitrary N-limb value (< 2^(32N)) below p by repeated subtraction (Fr: < 3r; Fp wire
// values are required to be canonical already, the loop then runs zero times)
template <class P> HD void fe_reduce_loose(Fe<P>& a) {
    constexpr int N = P::N;
    for (int it = 0; it < 8; it++) {
        uint32_t t[N];
        t[0] = sub_cc(a.l[0], P::p(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(a.l[i], P::p(i));
        uint32_t borrow = subc(0, 0);
        if (borrow) break;
#pragma unroll
        for (int i = 0; i < N; i++) a.l[i] = t[i];
    }
}

// a^(p-2) by square-and-multiply over the bits of p-2 (constant exponent)
template <class P> HD Fe<P> fe_inv(const Fe<P>& a) {
    constexpr int N = P::N;
    Fe<P> acc = fe_one<P>();
    for (int i = N * 32 - 1; i >= 0; i--) {
        acc = fe_sqr(acc);
        if ((P::pm2(i >> 5) >> (i & 31)) & 1) acc = fe_mul(acc, a);
    }
    return acc;
}
// Inverse by the binary extended Euclidean algorithm (shifts and subtractions only): about ten
// times shorter than the Fermat ladder on a single thread, which is what the serial tail of an MSM
// (one affine normalisation) pays for.  a is in Montgomery form and non-zero; so is the result.
template <class P> HD Fe<P> fe_inv_gcd(const Fe<P>& a) {
    constexpr int N = P::N;
    Fe<P> u = a, v, x1 = fe_zero<P>(), x2 = fe_zero<P>();   // invariants: x1*a == u, x2*a == v (mod p)
    for (int i = 0; i < N; i++) v.l[i] = P::p(i);
    x1.l[0] = 1;
    auto is_one = [](const Fe<P>& t) {
        uint32_t o = t.l[0] ^ 1u;
        for (int i = 1; i < N; i++) o |= t.l[i];
        return o == 0;
    };
    auto halve = [](Fe<P>& t, Fe<P>& x) {
        for (int i = 0; i < N - 1; i++) t.l[i] = (t.l[i] >> 1) | (t.l[i + 1] << 31);
        t.l[N - 1] >>= 1;
        if (x.l[0] & 1) {                                   // x odd: (x + p) / 2, x + p < 2^(32N)
            x.l[0] = add_cc(x.l[0], P::p(0));
            for (int i = 1; i < N - 1; i++) x.l[i] = addc_cc(x.l[i], P::p(i));
            x.l[N - 1] = addc(x.l[N - 1], P::p(N - 1));
        }
        for (int i = 0; i < N - 1; i++) x.l[i] = (x.l[i] >> 1) | (x.l[i + 1] << 31);
        x.l[N - 1] >>= 1;
    };
    for (int guard = 0; guard < 4 * 32 * N && !is_one(u) && !is_one(v); guard++) {
        while (!(u.l[0] & 1)) halve(u, x1);
        while (!(v.l[0] & 1)) halve(v, x2);
        // u >= v ?
        uint32_t t[N];
        t[0] = sub_cc(u.l[0], v.l[0]);
        for (int i = 1; i < N; i++) t[i] = subc_cc(u.l[i], v.l[i]);
        uint32_t borrow = subc(0, 0);
        if (!borrow) {
            for (int i = 0; i < N; i++) u.l[i] = t[i];
            x1 = fe_sub(x1, x2);
        } else {
            v.l[0] = sub_cc(v.l[0], u.l[0]);
            for (int i = 1; i < N - 1; i++) v.l[i] = subc_cc(v.l[i], u.l[i]);
            v.l[N - 1] = subc(v.l[N - 1], u.l[N - 1]);
            x2 = fe_sub(x2, x1);
        }
    }
    Fe<P> res = is_one(u) ? x1 : x2;                        // (aR)^-1 as a plain integer
    Fe<P> r2;
    for (int i = 0; i < N; i++) r2.l[i] = P::r2(i);
    return fe_mul(fe_mul(res, r2), r2);                     // * R^3 / R^2 = a^-1 R
}
// ---------------------------------------------------------------------------------------
// Branch-free modular inversion: the optimized binary GCD of T. Pornin ("Optimized Binary GCD for
// Modular Inversion", 2020).  30 GCD steps at a time are decided on 63-bit approximations of (a, b)
// (top 33 + low 30 bits) and applied to the full-width values as one linear combination with small
// signed factors; ceil((2*bits-1)/30) rounds.  No data-dependent branches, so the 32 lanes of a warp
// inverting 32 different values stay in lockstep -- which is what the batched-affine bucket
// accumulation needs -- and a single thread finishes in ~1/10 of the Fermat ladder.
// a is in Montgomery form and non-zero modulo p; so is the result.
// ---------------------------------------------------------------------------------------
namespace invdetail {
constexpr int K = 30;
// r (N+1 limbs, two's complement) = a*f + b*g for unsigned N-limb a, b and |f|, |g| <= 2^30
template <int N> HD void lincomb(uint32_t* r, const uint32_t* a, int64_t f, const uint32_t* b, int64_t g) {
    uint32_t fa = (uint32_t)(f < 0 ? -f : f), ga = (uint32_t)(g < 0 ? -g : g);
    uint32_t pa[N + 1], pb[N + 1];
    uint64_t ca = 0, cb = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        ca += (uint64_t)a[i] * fa;
        pa[i] = (uint32_t)ca;
        ca >>= 32;
        cb += (uint64_t)b[i] * ga;
        pb[i] = (uint32_t)cb;
        cb >>= 32;
    }
    pa[N] = (uint32_t)ca;
    pb[N] = (uint32_t)cb;
    // conditional negation by xor/add with the sign masks, then the sum
    uint32_t ma = f < 0 ? 0xffffffffu : 0u, mb = g < 0 ? 0xffffffffu : 0u;
    uint64_t c1 = ma & 1u, c2 = mb & 1u, c3 = 0;
#pragma unroll
    for (int i = 0; i <= N; i++) {
        c1 += (uint64_t)(pa[i] ^ ma);
        c2 += (uint64_t)(pb[i] ^ mb);
        c3 += (uint64_t)(uint32_t)c1 + (uint32_t)c2;
        r[i] = (uint32_t)c3;
        c1 >>= 32;
        c2 >>= 32;
        c3 >>= 32;
    }
}
// r (N+1 limbs, signed) >>= 30 arithmetically; result must fit N limbs plus sign
template <int N> HD void shr30(uint32_t* r) {
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = (r[i] >> K) | (r[i + 1] << (32 - K));
    r[N] = (uint32_t)((int32_t)r[N] >> K);
}
template <int N> HD void negate(uint32_t* r) {   // N+1 limbs
    uint64_t c = 1;
#pragma unroll
    for (int i = 0; i <= N; i++) {
        c += (uint64_t)(~r[i]);
        r[i] = (uint32_t)c;
        c >>= 32;
    }
}
}  // namespace invdetail

template <class P> HD Fe<P> fe_inv_fast(const Fe<P>& x) {
    using namespace invdetail;
    constexpr int N = P::N;
    constexpr int ROUNDS = (2 * 32 * N - 1 + K - 1) / K;     // safe for any value below 2^(32N)
    uint32_t a[N + 1], b[N + 1], u[N + 1], v[N + 1];
    for (int i = 0; i < N; i++) { a[i] = x.l[i]; b[i] = P::p(i); u[i] = 0; v[i] = 0; }
    a[N] = b[N] = u[N] = v[N] = 0;
    u[0] = 1;
    for (int round = 0; round < ROUNDS; round++) {
        // ---- 63-bit approximations of a and b (33 high bits, 30 exact low bits): exact if both are below 2^63
        uint32_t hi_a = 0, mid_a = 0, lo_a = 0, hi_b = 0, mid_b = 0, lo_b = 0;   // the three limbs under the top one
        bool found = false;
#pragma unroll
        for (int j = N - 1; j >= 2; j--) {
            bool here = !found && ((a[j] | b[j]) != 0);
            if (here) { hi_a = a[j]; mid_a = a[j - 1]; lo_a = a[j - 2]; hi_b = b[j]; mid_b = b[j - 1]; lo_b = b[j - 2]; }
            found = found || here;
        }
        uint64_t a_, b_;
        uint64_t xa = ((uint64_t)a[1] << 32) | a[0], xb = ((uint64_t)b[1] << 32) | b[0];
        if (!found && ((xa | xb) >> 63) == 0) {
            a_ = xa;
            b_ = xb;
        } else {
            if (!found) { hi_a = a[1]; mid_a = a[0]; lo_a = 0; hi_b = b[1]; mid_b = b[0]; lo_b = 0; }
            uint64_t ta = ((uint64_t)hi_a << 32) | mid_a, tb = ((uint64_t)hi_b << 32) | mid_b;
            uint64_t m = ta | tb;
            int sh = 0;                                       // leading zeros of m (m >= 2^32 here or top limb set)
            for (int t = 32; t >= 1; t >>= 1)
                if ((m >> (64 - t)) == 0) { m <<= t; sh += t; }
            if (sh) {                                         // 1 <= sh <= 31
                ta = (ta << sh) | ((uint64_t)lo_a >> (32 - sh));
                tb = (tb << sh) | ((uint64_t)lo_b >> (32 - sh));
            }
            a_ = ((ta >> 31) << K) | (a[0] & ((1u << K) - 1));
            b_ = ((tb >> 31) << K) | (b[0] & ((1u << K) - 1));
        }
        // ---- 30 binary-GCD steps on the approximations, recording the update factors
        int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
        for (int j = 0; j < K; j++) {
            uint64_t odd = (uint64_t)0 - (a_ & 1);
            uint64_t sw = odd & ((uint64_t)0 - (uint64_t)(a_ < b_));
            uint64_t t = (a_ ^ b_) & sw;
            a_ ^= t;
            b_ ^= t;
            int64_t tf = (f0 ^ f1) & (int64_t)sw, tg = (g0 ^ g1) & (int64_t)sw;
            f0 ^= tf; f1 ^= tf;
            g0 ^= tg; g1 ^= tg;
            a_ -= b_ & odd;
            f0 -= f1 & (int64_t)odd;
            g0 -= g1 & (int64_t)odd;
            a_ >>= 1;
            f1 <<= 1;
            g1 <<= 1;
        }
        // ---- apply to (a, b): exact division by 2^30, then fix the signs
        uint32_t na[N + 1], nb[N + 1];
        lincomb<N>(na, a, f0, b, g0);
        lincomb<N>(nb, a, f1, b, g1);
        shr30<N>(na);
        shr30<N>(nb);
        if ((int32_t)na[N] < 0) { negate<N>(na); f0 = -f0; g0 = -g0; }
        if ((int32_t)nb[N] < 0) { negate<N>(nb); f1 = -f1; g1 = -g1; }
        for (int i = 0; i <= N; i++) { a[i] = na[i]; b[i] = nb[i]; }
        // ---- apply to (u, v) modulo p: add the multiple of p that clears the low 30 bits, shift,
        //      and bring the result from (-p, 2p) back into [0, p)
        uint32_t nu[N + 1], nv[N + 1];
        lincomb<N>(nu, u, f0, v, g0);
        lincomb<N>(nv, u, f1, v, g1);
#pragma unroll
        for (int which = 0; which < 2; which++) {
            uint32_t* t = which ? nv : nu;
            uint32_t q = (t[0] * P::M0) & ((1u << K) - 1);
            uint64_t c = 0;
            // t += q * p   (sign-extended add over N+1 limbs)
#pragma unroll
            for (int i = 0; i < N; i++) {
                c += (uint64_t)q * P::p(i) + t[i];
                t[i] = (uint32_t)c;
                c >>= 32;
            }
            t[N] = t[N] + (uint32_t)c;
            shr30<N>(t);
            if ((int32_t)t[N] < 0) {                           // t += p
                uint64_t cc = 0;
#pragma unroll
                for (int i = 0; i < N; i++) { cc += (uint64_t)t[i] + P::p(i); t[i] = (uint32_t)cc; cc >>= 32; }
                t[N] += (uint32_t)cc;
            }
            // t >= p ?  (t is now non-negative and below 2p)
            uint32_t d[N];
            uint64_t bw = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                uint64_t s = (uint64_t)t[i] - P::p(i) - bw;
                d[i] = (uint32_t)s;
                bw = (s >> 32) & 1;
            }
            if (!bw) {
#pragma unroll
                for (int i = 0; i < N; i++) t[i] = d[i];
            }
            t[N] = 0;
        }
        for (int i = 0; i <= N; i++) { u[i] = nu[i]; v[i] = nv[i]; }
    }
    // b == 1 and v == (xR)^-1 as a plain integer; back to Montgomery form: * R^3 / R^2
    Fe<P> res, r2;
    for (int i = 0; i < N; i++) { res.l[i] = v[i]; r2.l[i] = P::r2(i); }
    return fe_mul(fe_mul(res, r2), r2);
}

// a^e for a small runtime exponent
template <class P> HD Fe<P> fe_pow_u64(const Fe<P>& a, uint64_t e) {
    Fe<P> acc = fe_one<P>(), base = a;
    while (e) {
        if (e & 1) acc = fe_mul(acc, base);
        base = fe_sqr(base);
        e >>= 1;
    }
    return acc;
}

// Fr with all eight m*r products on the multiplier pipe (experiment: is the NTT bound by the FMA or by the ALU pipe?)
struct FrParamsPlain : FrParams {
    static constexpr bool LOW_LIMBS_TRIVIAL = false;
};
typedef Fe<FpParams> Fp;
typedef Fe<FrParams> Fr;

}  // namespace b200zk
