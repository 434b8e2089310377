// srs.cuh -- G1 kernels next to the MSM: SRS generation by fixed-base multiplication and batched
// decompression of proof points.
//
//   * ParamsKZG::unsafe_setup (/root/reference/src/kzg_params.rs:33-80 calls it and caches the result):
//     g[i] = s^i * G and g_lagrange[i] = L_i(s) * G with L_i(s) = omega^i (s^n - 1) / (n (s - omega^i)).
//     Both tables are n fixed-base multiplications of 255-bit scalars: 32 byte-windows against a table of
//     32 x 255 multiples of G (786 KB, L2-resident), 32 mixed additions per point instead of ~380
//     double-and-add steps, then one in-thread affine normalisation.
//   * the verifier's front end: every proof carries its commitments as 48-byte ZCash-compressed points
//     (/root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:62-83, bls_utils.ak:17-49).
#pragma once
#include "g1_call.cuh"

namespace b200zk {

// (p + 1) / 4, the square-root exponent (p = 3 mod 4)
__device__ __constant__ uint32_t FP_SQRT_EXP_WORDS[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                                                          0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};

// ---------------------------------------------------------------------------------------
// batched G1 decompression: the front end of batch verification (every proof carries its
// commitments as 48-byte ZCash-compressed points,
// /root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:62-83, bls_utils.ak:17-49).
// One thread per point: y = (x^3 + 4)^((p+1)/4) (p = 3 mod 4,
// /root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/CompressUncompress.hs:98),
// sign chosen by the "y is the larger root" flag.  status[i]: 0 ok, 1 not a compressed encoding,
// 2 bad infinity encoding, 3 x >= p, 4 x not on the curve.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) g1_decompress_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ out_canon,
                                                            uint32_t* __restrict__ out_mont, uint32_t* __restrict__ status) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* c = in + 48 * i;
    uint32_t flags = c[0] & 0xE0u;
    Fp x;                                              // big-endian bytes -> little-endian limbs
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const uint8_t* q = c + 44 - 4 * k;
        x.l[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    x.l[11] &= 0x1fffffffu;
    Fp zero = fe_zero<FpParams>();
    uint32_t st = 0;
    Fp xo = zero, yo = zero;
    if (!(flags & 0x80u)) st = 1;
    else if (flags & 0x40u) {
        if (!fe_is_zero(x) || (flags & 0x20u)) st = 2;
    } else {
        Fp t = x;
        fe_reduce_loose(t);
        if (!fe_eq(t, x)) st = 3;
        else {
            Fp r2;
            for (int k = 0; k < 12; k++) r2.l[k] = FpParams::r2(k);
            Fp xm = fp_mul_ni(x, r2);                           // to Montgomery form
            Fp rhs = fe_add(fp_mul_ni(fp_mul_ni(xm, xm), xm), fe_dbl(fe_dbl(fe_one<FpParams>())));
            Fp y = fe_one<FpParams>();                          // rhs^((p+1)/4), left-to-right
            for (int b = 380; b >= 0; b--) {
                y = fp_mul_ni(y, y);
                if ((FP_SQRT_EXP_WORDS[b >> 5] >> (b & 31)) & 1) y = fp_mul_ni(y, rhs);
            }
            if (!fe_eq(fp_mul_ni(y, y), rhs)) st = 4;
            else {
                Fp yc = fe_from_mont(y);                        // canonical y; larger root iff 2y > p
                Fp ny = fe_sub(zero, yc);                       // p - y as a canonical integer (y != 0)
                bool larger = false, decided = false;
#pragma unroll
                for (int k = 11; k >= 0; k--) {
                    if (!decided && yc.l[k] != ny.l[k]) { larger = yc.l[k] > ny.l[k]; decided = true; }
                }
                if (larger != ((flags & 0x20u) != 0)) { yc = ny; y = fe_neg(y); }
                xo = x; yo = yc;
                if (out_mont) { fp_st(out_mont + 24 * i, xm); fp_st(out_mont + 24 * i + 12, y); }
            }
        }
    }
    if (st != 0 || (flags & 0x40u)) {
        if (out_mont) { fp_st(out_mont + 24 * i, zero); fp_st(out_mont + 24 * i + 12, zero); }
    }
    if (out_canon) { fp_st(out_canon + 24 * i, xo); fp_st(out_canon + 24 * i + 12, yo); }
    status[i] = st;
}


// ---------------------------------------------------------------------------------------
// fixed-base multiplication of 255-bit scalars
// ---------------------------------------------------------------------------------------
constexpr int FIXED_WINDOWS = 32;   // byte windows
// table[w*256 + d] = d * 2^(8w) * G (affine Montgomery; entry d = 0 unused), one thread per entry
__global__ void __launch_bounds__(128) g1_fixed_table32_kernel(const uint32_t* __restrict__ gen_mont, uint32_t* __restrict__ table) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= FIXED_WINDOWS * 256) return;
    uint32_t w = t >> 8, d = t & 255;
    G1Affine g = g1a_ldg(gen_mont, 0);
    G1Xyzz acc;
    xyzz_set_inf(acc);
    for (int b = 7; b >= 0; b--) {
        xyzz_dbl_ni(acc);
        if ((d >> b) & 1) xyzz_add_mixed_ni(acc, g, false);
    }
    for (uint32_t k = 0; k < 8 * w; k++) xyzz_dbl_ni(acc);
    G1Affine a = xyzz_to_affine_ni(acc);
    fp_st(table + 24 * t, a.x);
    fp_st(table + 24 * t + 12, a.y);
}
// out[i] = scalars[i] * G, packed Montgomery affine; scalars are Fr in Montgomery (fmt_mont) or canonical form
__global__ void __launch_bounds__(128) g1_fixed_mul_kernel(const uint32_t* __restrict__ table, const uint32_t* __restrict__ scalars,
                                                           uint32_t fmt_mont, uint64_t n, uint32_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(scalars + 8 * i);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
    v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    if (fmt_mont) v = fe_from_mont(v);
    else fe_reduce_loose(v);
    G1Xyzz acc;
    xyzz_set_inf(acc);
    for (int w = 0; w < FIXED_WINDOWS; w++) {
        uint32_t d = (v.l[w >> 2] >> (8 * (w & 3))) & 255;
        if (d) xyzz_add_mixed_ni(acc, g1a_ldg(table, (uint64_t)w * 256 + d), false);
    }
    G1Affine r = xyzz_to_affine_ni(acc);
    fp_st(out + 24 * i, r.x);
    fp_st(out + 24 * i + 12, r.y);
}
// packed Montgomery affine -> canonical wire format (x || y little-endian integers), for host-buffer output
__global__ void __launch_bounds__(128) g1_to_canonical_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_st(out + 24 * i, fe_from_mont(fp_ld(in + 24 * i)));
    fp_st(out + 24 * i + 12, fe_from_mont(fp_ld(in + 24 * i + 12)));
}

// Lagrange-basis scalars of the SRS, two sweeps around a batched inversion:
//   pre : den[i] = s - omega^i                      (flag set if some den is zero: s lies in the domain)
//   post: out[i] = den[i]^-1 * omega^i * c,  c = (s^n - 1) / n
__global__ void __launch_bounds__(256) srs_lagrange_pre_kernel(const uint32_t* __restrict__ omega_pows, const uint32_t* __restrict__ s_p,
                                                               uint32_t* __restrict__ den, uint64_t n, uint32_t* __restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(omega_pows + 8 * i);
    uint4 a = q[0], b = q[1];
    Fr w, s;
    w.l[0] = a.x; w.l[1] = a.y; w.l[2] = a.z; w.l[3] = a.w; w.l[4] = b.x; w.l[5] = b.y; w.l[6] = b.z; w.l[7] = b.w;
    for (int k = 0; k < 8; k++) s.l[k] = s_p[k];
    Fr d = fe_sub(s, w);
    if (fe_is_zero(d)) atomicOr(flag, 1u);
    uint4* o = reinterpret_cast<uint4*>(den + 8 * i);
    o[0] = make_uint4(d.l[0], d.l[1], d.l[2], d.l[3]);
    o[1] = make_uint4(d.l[4], d.l[5], d.l[6], d.l[7]);
}

}  // namespace b200zk
