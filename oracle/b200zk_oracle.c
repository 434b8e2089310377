/*
 * b200zk_oracle.c -- CPU restatement (plain C, 64-bit limbs, unsigned __int128)
 * of the hot path under the Halo2/KZG prover: BLS12-381 G1 multi-scalar
 * multiplication and the Fr NTT of the halo2 EvaluationDomain.
 *
 * TEST INFRASTRUCTURE ONLY.  The product (plutus-halo2-verifier-gen_b200/) never
 * links, loads or calls this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may, and only as the checker
 * or as the timed CPU baseline.
 *
 * Parity status.  The reference's own arithmetic for this path lives in crates
 * that are not under /root/reference (midnight-proofs =0.8.0, midnight-curves
 * =0.3.0 -> blst; /root/reference/Cargo.toml:27-28) and there is no Rust
 * toolchain, so oracle/_ref cannot be built.  This file restates the published
 * algorithms and is pinned (tests/test_oracle_kats.py) against every
 * known-answer vector the reference's tests hold for the path:
 *   field moduli          plinth-verifier/plutus-halo2/src/Plutus/Crypto/BlsTypes.hs:97,102-103
 *   generator 7 / DELTA   plinth-verifier/plutus-halo2/src/Plutus/Crypto/Constants.hs:10-13
 *   omega_k convention    aiken-verifier/aiken_halo2/lib/omega_rotations.ak:48-81
 *   G1 compressed format  aiken-verifier/aiken_halo2/lib/bls_utils.ak:17-49,
 *                         aiken-verifier/aiken_halo2/lib/transcript.ak:121-156 (G, -G, 42*G)
 *   Fr wire format        aiken-verifier/aiken_halo2/lib/transcript.ak:29-45,158-179
 *   golden proof points   aiken-verifier/aiken_halo2/lib/transcript.ak:241-382
 * An MSM of size > 1 and an NTT output vector are pinned by NO reference test
 * ("parity unpinned" for those); they rest on the uniqueness of the group
 * element / DFT vector and on cross-checks inside this oracle (double-and-add
 * vs Pippenger, O(n^2) DFT vs radix-2) and against oracle/pyref.py.
 *
 * Reference call sites this restates (the upstream functions they reach):
 *   commit / commit_lagrange (MSM)  examples/simple_mul.rs:62,72; src/circuits/atms_circuit.rs:247,296
 *   EvaluationDomain FFTs           examples/ivc.rs:109; src/circuits/ivc_circuit.rs:305
 *   verifier final MSM              aiken-verifier/aiken_halo2/lib/halo2_kzg.ak:15-44,
 *                                   plinth-verifier/.../Halo2/MSMEval.hs:20-25
 *
 * Wire formats (same as include/b200zk.h): Fr = 32 B little-endian canonical;
 * G1 affine = x||y, 48 B little-endian canonical each, identity = (0,0)
 * (plinth-verifier/.../Halo2/CompressUncompress.hs:72).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

#define FPN 6
#define FRN 4
#define INLINE static inline __attribute__((always_inline))

/* ------------------------------------------------------------------ moduli */
static const u64 FP_P[FPN] = {
    0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
    0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const u64 FR_P[FRN] = {
    0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
    0x73eda753299d7d48ULL};

static u64 FP_INV, FR_INV;           /* -p^-1 mod 2^64 */
static u64 FP_R[FPN], FP_R2[FPN];    /* R mod p, R^2 mod p */
static u64 FR_R[FRN], FR_R2[FRN];
static int g_init_done = 0;

/* ------------------------------------------------------------------ generic n-limb helpers */
INLINE int ge_n(const u64 *a, const u64 *b, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
INLINE u64 add_n(u64 *r, const u64 *a, const u64 *b, int n) {
    u64 c = 0;
    for (int i = 0; i < n; i++) {
        u128 t = (u128)a[i] + b[i] + c;
        r[i] = (u64)t;
        c = (u64)(t >> 64);
    }
    return c;
}
INLINE u64 sub_n(u64 *r, const u64 *a, const u64 *b, int n) {
    u64 br = 0;
    for (int i = 0; i < n; i++) {
        u128 t = (u128)a[i] - b[i] - br;
        r[i] = (u64)t;
        br = (u64)(t >> 64) & 1;
    }
    return br;
}
INLINE int is_zero_n(const u64 *a, int n) {
    u64 o = 0;
    for (int i = 0; i < n; i++) o |= a[i];
    return o == 0;
}
INLINE void mod_add_n(u64 *r, const u64 *a, const u64 *b, const u64 *p, int n) {
    u64 t[FPN];
    u64 c = add_n(t, a, b, n);
    if (c || ge_n(t, p, n)) sub_n(t, t, p, n);
    memcpy(r, t, 8 * n);
}
INLINE void mod_sub_n(u64 *r, const u64 *a, const u64 *b, const u64 *p, int n) {
    u64 t[FPN];
    if (sub_n(t, a, b, n)) add_n(t, t, p, n);
    memcpy(r, t, 8 * n);
}
/* CIOS Montgomery multiplication, r = a*b/2^(64n) mod p, inputs < p */
INLINE void mont_mul_n(u64 *r, const u64 *a, const u64 *b, const u64 *p, u64 inv, int n) {
    u64 t[FPN + 2];
    for (int i = 0; i < n + 2; i++) t[i] = 0;
    for (int i = 0; i < n; i++) {
        u64 c = 0;
        for (int j = 0; j < n; j++) {
            u128 s = (u128)a[j] * b[i] + t[j] + c;
            t[j] = (u64)s;
            c = (u64)(s >> 64);
        }
        u128 s = (u128)t[n] + c;
        t[n] = (u64)s;
        t[n + 1] = (u64)(s >> 64);
        u64 m = t[0] * inv;
        s = (u128)m * p[0] + t[0];
        c = (u64)(s >> 64);
        for (int j = 1; j < n; j++) {
            s = (u128)m * p[j] + t[j] + c;
            t[j - 1] = (u64)s;
            c = (u64)(s >> 64);
        }
        s = (u128)t[n] + c;
        t[n - 1] = (u64)s;
        t[n] = t[n + 1] + (u64)(s >> 64);
    }
    if (t[n] || ge_n(t, p, n)) sub_n(t, t, p, n);
    memcpy(r, t, 8 * n);
}

/* ------------------------------------------------------------------ Fp (381-bit) */
typedef struct { u64 l[FPN]; } fp;
typedef struct { u64 l[FRN]; } fr;

INLINE void fp_mul(fp *r, const fp *a, const fp *b) { mont_mul_n(r->l, a->l, b->l, FP_P, FP_INV, FPN); }
INLINE void fp_sqr(fp *r, const fp *a) { mont_mul_n(r->l, a->l, a->l, FP_P, FP_INV, FPN); }
INLINE void fp_add(fp *r, const fp *a, const fp *b) { mod_add_n(r->l, a->l, b->l, FP_P, FPN); }
INLINE void fp_sub(fp *r, const fp *a, const fp *b) { mod_sub_n(r->l, a->l, b->l, FP_P, FPN); }
INLINE int fp_is_zero(const fp *a) { return is_zero_n(a->l, FPN); }
INLINE int fp_eq(const fp *a, const fp *b) { return memcmp(a->l, b->l, 8 * FPN) == 0; }
INLINE void fp_neg(fp *r, const fp *a) {
    if (fp_is_zero(a)) { *r = *a; return; }
    sub_n(r->l, FP_P, a->l, FPN);
}
INLINE void fr_mul(fr *r, const fr *a, const fr *b) { mont_mul_n(r->l, a->l, b->l, FR_P, FR_INV, FRN); }
INLINE void fr_add(fr *r, const fr *a, const fr *b) { mod_add_n(r->l, a->l, b->l, FR_P, FRN); }
INLINE void fr_sub(fr *r, const fr *a, const fr *b) { mod_sub_n(r->l, a->l, b->l, FR_P, FRN); }

static void fp_from_bytes(fp *r, const uint8_t b[48]) { /* canonical LE -> Montgomery */
    fp t, r2;
    memcpy(t.l, b, 48);
    memcpy(r2.l, FP_R2, 48);
    fp_mul(r, &t, &r2);
}
static void fp_to_bytes(uint8_t b[48], const fp *a) { /* Montgomery -> canonical LE */
    fp one = {{1, 0, 0, 0, 0, 0}}, t;
    fp_mul(&t, a, &one);
    memcpy(b, t.l, 48);
}
static void fr_from_bytes(fr *r, const uint8_t b[32]) {
    fr t, r2;
    memcpy(t.l, b, 32);
    /* reduce on read: value < 2^256 < 3r, so at most two subtractions (transcript.ak:158-179) */
    while (ge_n(t.l, FR_P, FRN)) sub_n(t.l, t.l, FR_P, FRN);
    memcpy(r2.l, FR_R2, 32);
    fr_mul(r, &t, &r2);
}
static void fr_to_bytes(uint8_t b[32], const fr *a) {
    fr one = {{1, 0, 0, 0}}, t;
    fr_mul(&t, a, &one);
    memcpy(b, t.l, 32);
}

/* a^e, e given as little-endian limbs (not Montgomery) */
static void fp_pow(fp *r, const fp *a, const u64 *e, int en) {
    fp acc, base = *a;
    memcpy(acc.l, FP_R, 48);
    for (int i = 0; i < en * 64; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) fp_mul(&acc, &acc, &base);
        fp_sqr(&base, &base);
    }
    *r = acc;
}
static void fp_inv(fp *r, const fp *a) {
    u64 e[FPN], two[FPN] = {2, 0, 0, 0, 0, 0};
    sub_n(e, FP_P, two, FPN);
    fp_pow(r, a, e, FPN);
}
static void fr_pow(fr *r, const fr *a, const u64 *e, int en) {
    fr acc, base = *a;
    memcpy(acc.l, FR_R, 32);
    for (int i = 0; i < en * 64; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) fr_mul(&acc, &acc, &base);
        fr_mul(&base, &base, &base);
    }
    *r = acc;
}
static void fr_inv(fr *r, const fr *a) {
    u64 e[FRN], two[FRN] = {2, 0, 0, 0};
    sub_n(e, FR_P, two, FRN);
    fr_pow(r, a, e, FRN);
}

static u64 neg_inv64(u64 p0) {
    u64 x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p0 * x; /* Newton: x = p0^-1 mod 2^64 */
    return (u64)0 - x;
}
static void pow2_mod(u64 *r, int bits, const u64 *p, int n) { /* 2^bits mod p */
    u64 t[FPN] = {0};
    t[0] = 1;
    for (int i = 0; i < bits; i++) mod_add_n(t, t, t, p, n);
    memcpy(r, t, 8 * n);
}

void orc_init(void) {
    if (g_init_done) return;
    FP_INV = neg_inv64(FP_P[0]);
    FR_INV = neg_inv64(FR_P[0]);
    pow2_mod(FP_R, 384, FP_P, FPN);
    pow2_mod(FP_R2, 768, FP_P, FPN);
    pow2_mod(FR_R, 256, FR_P, FRN);
    pow2_mod(FR_R2, 512, FR_P, FRN);
    g_init_done = 1;
}

/* ------------------------------------------------------------------ exported field ops (canonical bytes) */
void orc_fr_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    fr x, y; orc_init(); fr_from_bytes(&x, a); fr_from_bytes(&y, b); fr_mul(&x, &x, &y); fr_to_bytes(out, &x);
}
void orc_fr_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    fr x, y; orc_init(); fr_from_bytes(&x, a); fr_from_bytes(&y, b); fr_add(&x, &x, &y); fr_to_bytes(out, &x);
}
void orc_fr_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    fr x, y; orc_init(); fr_from_bytes(&x, a); fr_from_bytes(&y, b); fr_sub(&x, &x, &y); fr_to_bytes(out, &x);
}
void orc_fr_inv(const uint8_t a[32], uint8_t out[32]) {
    fr x; orc_init(); fr_from_bytes(&x, a); fr_inv(&x, &x); fr_to_bytes(out, &x);
}
/* out = a * 2^256 mod r as raw limbs: the in-memory (Montgomery) form of a blst_fr */
void orc_fr_to_mont(const uint8_t a[32], uint8_t out[32]) {
    fr x; orc_init(); fr_from_bytes(&x, a); memcpy(out, x.l, 32);
}
void orc_fp_mul(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]) {
    fp x, y; orc_init(); fp_from_bytes(&x, a); fp_from_bytes(&y, b); fp_mul(&x, &x, &y); fp_to_bytes(out, &x);
}
void orc_fp_add(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]) {
    fp x, y; orc_init(); fp_from_bytes(&x, a); fp_from_bytes(&y, b); fp_add(&x, &x, &y); fp_to_bytes(out, &x);
}
void orc_fp_sub(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]) {
    fp x, y; orc_init(); fp_from_bytes(&x, a); fp_from_bytes(&y, b); fp_sub(&x, &x, &y); fp_to_bytes(out, &x);
}
void orc_fp_inv(const uint8_t a[48], uint8_t out[48]) {
    fp x; orc_init(); fp_from_bytes(&x, a); fp_inv(&x, &x); fp_to_bytes(out, &x);
}
/* out = a * 2^384 mod p as raw limbs: the in-memory (Montgomery) form of a blst_fp */
void orc_fp_to_mont(const uint8_t a[48], uint8_t out[48]) {
    fp x; orc_init(); fp_from_bytes(&x, a); memcpy(out, x.l, 48);
}

/* ------------------------------------------------------------------ G1: Jacobian (X,Y,Z), Z=0 is infinity */
typedef struct { fp x, y; int inf; } g1a;
typedef struct { fp x, y, z; } g1j;

static void g1a_from_wire(g1a *r, const uint8_t b[96]) {
    int z = 1;
    for (int i = 0; i < 96; i++) if (b[i]) { z = 0; break; }
    r->inf = z;
    fp_from_bytes(&r->x, b);
    fp_from_bytes(&r->y, b + 48);
}
static void g1j_set_inf(g1j *r) { memset(r, 0, sizeof(*r)); }
static int g1j_is_inf(const g1j *a) { return fp_is_zero(&a->z); }
static void g1j_from_affine(g1j *r, const g1a *a) {
    if (a->inf) { g1j_set_inf(r); return; }
    r->x = a->x; r->y = a->y; memcpy(r->z.l, FP_R, 48);
}
/* dbl-2009-l (a = 0) */
static void g1j_double(g1j *r, const g1j *p) {
    if (g1j_is_inf(p)) { *r = *p; return; }
    fp A, B, C, D, E, F, t;
    fp_sqr(&A, &p->x);
    fp_sqr(&B, &p->y);
    fp_sqr(&C, &B);
    fp_add(&t, &p->x, &B); fp_sqr(&t, &t); fp_sub(&t, &t, &A); fp_sub(&t, &t, &C); fp_add(&D, &t, &t);
    fp_add(&E, &A, &A); fp_add(&E, &E, &A);
    fp_sqr(&F, &E);
    fp z3; fp_mul(&z3, &p->y, &p->z); fp_add(&z3, &z3, &z3);
    fp x3; fp_sub(&x3, &F, &D); fp_sub(&x3, &x3, &D);
    fp c8; fp_add(&c8, &C, &C); fp_add(&c8, &c8, &c8); fp_add(&c8, &c8, &c8);
    fp y3; fp_sub(&y3, &D, &x3); fp_mul(&y3, &E, &y3); fp_sub(&y3, &y3, &c8);
    r->x = x3; r->y = y3; r->z = z3;
}
/* general Jacobian add (add-2007-bl shape, with the equal / opposite cases handled) */
static void g1j_add(g1j *r, const g1j *p, const g1j *q) {
    if (g1j_is_inf(p)) { *r = *q; return; }
    if (g1j_is_inf(q)) { *r = *p; return; }
    fp z1z1, z2z2, u1, u2, s1, s2, h, rr, t;
    fp_sqr(&z1z1, &p->z); fp_sqr(&z2z2, &q->z);
    fp_mul(&u1, &p->x, &z2z2); fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&t, &q->z, &z2z2); fp_mul(&s1, &p->y, &t);
    fp_mul(&t, &p->z, &z1z1); fp_mul(&s2, &q->y, &t);
    fp_sub(&h, &u2, &u1); fp_sub(&rr, &s2, &s1);
    if (fp_is_zero(&h)) {
        if (fp_is_zero(&rr)) { g1j_double(r, p); return; }
        g1j_set_inf(r); return;
    }
    fp hh, hhh, v;
    fp_sqr(&hh, &h); fp_mul(&hhh, &hh, &h); fp_mul(&v, &u1, &hh);
    fp x3; fp_sqr(&x3, &rr); fp_sub(&x3, &x3, &hhh); fp_sub(&x3, &x3, &v); fp_sub(&x3, &x3, &v);
    fp y3; fp_sub(&y3, &v, &x3); fp_mul(&y3, &rr, &y3); fp_mul(&t, &s1, &hhh); fp_sub(&y3, &y3, &t);
    fp z3; fp_mul(&z3, &p->z, &q->z); fp_mul(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
/* mixed add, q affine; neg != 0 adds -q */
static void g1j_add_affine(g1j *r, const g1j *p, const g1a *q, int neg) {
    if (q->inf) { *r = *p; return; }
    fp qy = q->y;
    if (neg) fp_neg(&qy, &qy);
    if (g1j_is_inf(p)) { r->x = q->x; r->y = qy; memcpy(r->z.l, FP_R, 48); return; }
    fp z1z1, u2, s2, h, rr, t;
    fp_sqr(&z1z1, &p->z);
    fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&t, &p->z, &z1z1); fp_mul(&s2, &qy, &t);
    fp_sub(&h, &u2, &p->x); fp_sub(&rr, &s2, &p->y);
    if (fp_is_zero(&h)) {
        if (fp_is_zero(&rr)) { g1j_double(r, p); return; }
        g1j_set_inf(r); return;
    }
    fp hh, hhh, v;
    fp_sqr(&hh, &h); fp_mul(&hhh, &hh, &h); fp_mul(&v, &p->x, &hh);
    fp x3; fp_sqr(&x3, &rr); fp_sub(&x3, &x3, &hhh); fp_sub(&x3, &x3, &v); fp_sub(&x3, &x3, &v);
    fp y3; fp_sub(&y3, &v, &x3); fp_mul(&y3, &rr, &y3); fp_mul(&t, &p->y, &hhh); fp_sub(&y3, &y3, &t);
    fp z3; fp_mul(&z3, &p->z, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_to_wire(uint8_t out[96], const g1j *p) {
    if (g1j_is_inf(p)) { memset(out, 0, 96); return; }
    fp zi, zi2, zi3, x, y;
    fp_inv(&zi, &p->z); fp_sqr(&zi2, &zi); fp_mul(&zi3, &zi2, &zi);
    fp_mul(&x, &p->x, &zi2); fp_mul(&y, &p->y, &zi3);
    fp_to_bytes(out, &x); fp_to_bytes(out + 48, &y);
}
static void g1j_mul_bits(g1j *r, const g1a *p, const u64 *k, int nbits) {
    g1j acc; g1j_set_inf(&acc);
    for (int i = nbits - 1; i >= 0; i--) {
        g1j_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) g1j_add_affine(&acc, &acc, p, 0);
    }
    *r = acc;
}
static void scalar_canon(u64 k[FRN], const uint8_t s[32]) {
    memcpy(k, s, 32);
    while (ge_n(k, FR_P, FRN)) sub_n(k, k, FR_P, FRN);
}

static const uint8_t G1_GEN_WIRE_X_BE[48] = {
    0x17,0xf1,0xd3,0xa7,0x31,0x97,0xd7,0x94,0x26,0x95,0x63,0x8c,0x4f,0xa9,0xac,0x0f,
    0xc3,0x68,0x8c,0x4f,0x97,0x74,0xb9,0x05,0xa1,0x4e,0x3a,0x3f,0x17,0x1b,0xac,0x58,
    0x6c,0x55,0xe8,0x3f,0xf9,0x7a,0x1a,0xef,0xfb,0x3a,0xf0,0x0a,0xdb,0x22,0xc6,0xbb};
static const uint8_t G1_GEN_WIRE_Y_BE[48] = {
    0x08,0xb3,0xf4,0x81,0xe3,0xaa,0xa0,0xf1,0xa0,0x9e,0x30,0xed,0x74,0x1d,0x8a,0xe4,
    0xfc,0xf5,0xe0,0x95,0xd5,0xd0,0x0a,0xf6,0x00,0xdb,0x18,0xcb,0x2c,0x04,0xb3,0xed,
    0xd0,0x3c,0xc7,0x44,0xa2,0x88,0x8a,0xe4,0x0c,0xaa,0x23,0x29,0x46,0xc5,0xe7,0xe1};

void orc_g1_generator(uint8_t out[96]) {
    for (int i = 0; i < 48; i++) { out[i] = G1_GEN_WIRE_X_BE[47 - i]; out[48 + i] = G1_GEN_WIRE_Y_BE[47 - i]; }
}

int orc_g1_on_curve(const uint8_t p[96]) {
    orc_init();
    g1a a; g1a_from_wire(&a, p);
    if (a.inf) return 1;
    fp y2, x3, four, t;
    fp_sqr(&y2, &a.y); fp_sqr(&x3, &a.x); fp_mul(&x3, &x3, &a.x);
    memcpy(t.l, FP_R, 48); fp_add(&four, &t, &t); fp_add(&four, &four, &four);
    fp_add(&x3, &x3, &four);
    return fp_eq(&y2, &x3);
}
void orc_g1_add(const uint8_t a[96], const uint8_t b[96], uint8_t out[96]) {
    orc_init();
    g1a pa, pb; g1j j;
    g1a_from_wire(&pa, a); g1a_from_wire(&pb, b);
    g1j_from_affine(&j, &pa); g1j_add_affine(&j, &j, &pb, 0);
    g1j_to_wire(out, &j);
}
void orc_g1_mul(const uint8_t p[96], const uint8_t s[32], uint8_t out[96]) {
    orc_init();
    g1a a; g1j j; u64 k[FRN];
    g1a_from_wire(&a, p); scalar_canon(k, s);
    g1j_mul_bits(&j, &a, k, 255);
    g1j_to_wire(out, &j);
}
/* ZCash compressed encoding (bls_utils.ak:17-49): big-endian x, flags 0x80|0x40 inf|0x20 y larger */
void orc_g1_compress(const uint8_t p[96], uint8_t out[48]) {
    int z = 1;
    for (int i = 0; i < 96; i++) if (p[i]) { z = 0; break; }
    if (z) { memset(out, 0, 48); out[0] = 0xC0; return; }
    for (int i = 0; i < 48; i++) out[i] = p[47 - i];
    /* y > p - y  <=>  2y > p */
    u64 y[FPN], ny[FPN];
    memcpy(y, p + 48, 48);
    sub_n(ny, FP_P, y, FPN);
    int larger = ge_n(y, ny, FPN) && memcmp(y, ny, 48) != 0;
    out[0] |= 0x80 | (larger ? 0x20 : 0);
}
int orc_g1_decompress(const uint8_t in[48], uint8_t out[96]) {
    orc_init();
    if (!(in[0] & 0x80)) return -1;
    uint8_t xb[48];
    for (int i = 0; i < 48; i++) xb[i] = in[47 - i];
    xb[47] &= 0x1F;
    if (in[0] & 0x40) {
        for (int i = 0; i < 48; i++) if (xb[i]) return -2;
        if (in[0] & 0x20) return -2;
        memset(out, 0, 96); return 0;
    }
    u64 xl[FPN]; memcpy(xl, xb, 48);
    if (ge_n(xl, FP_P, FPN)) return -3;
    fp x, t, four, y;
    fp_from_bytes(&x, xb);
    fp_sqr(&t, &x); fp_mul(&t, &t, &x);
    memcpy(four.l, FP_R, 48); fp_add(&four, &four, &four); fp_add(&four, &four, &four);
    fp_add(&t, &t, &four);
    u64 e[FPN], one[FPN] = {1, 0, 0, 0, 0, 0};
    add_n(e, FP_P, one, FPN);                 /* (p+1)/4, p+1 does not overflow 384 bits */
    for (int i = 0; i < FPN; i++) e[i] = (e[i] >> 2) | (i + 1 < FPN ? e[i + 1] << 62 : 0);
    fp_pow(&y, &t, e, FPN);
    fp chk; fp_sqr(&chk, &y);
    if (!fp_eq(&chk, &t)) return -4;
    uint8_t yb[48]; fp_to_bytes(yb, &y);
    u64 yl[FPN], nyl[FPN]; memcpy(yl, yb, 48); sub_n(nyl, FP_P, yl, FPN);
    int larger = ge_n(yl, nyl, FPN) && memcmp(yl, nyl, 48) != 0;
    if (larger != !!(in[0] & 0x20)) memcpy(yb, nyl, 48);
    memcpy(out, xb, 48); memcpy(out + 48, yb, 48);
    return 0;
}

/* naive MSM: sum of double-and-add scalar multiplications */
void orc_g1_msm_naive(const uint8_t *points, const uint8_t *scalars, u64 n, uint8_t out[96]) {
    orc_init();
    g1j acc; g1j_set_inf(&acc);
    for (u64 i = 0; i < n; i++) {
        g1a a; g1j j; u64 k[FRN];
        g1a_from_wire(&a, points + 96 * i); scalar_canon(k, scalars + 32 * i);
        g1j_mul_bits(&j, &a, k, 255);
        g1j_add(&acc, &acc, &j);
    }
    g1j_to_wire(out, &acc);
}

/* Pippenger bucket method, signed c-bit digits, windows processed in parallel.
 * Restates the published algorithm that blst's p1s_mult_pippenger implements (the crate
 * is not under /root/reference); any correct MSM returns the same group element. */
static int pick_window(u64 n) {
    int c = 3;
    while (c < 16 && ((u64)1 << (c + 3)) <= n) c++;   /* roughly log2(n) - 3 */
    return c;
}
void orc_g1_msm(const uint8_t *points, const uint8_t *scalars, u64 n, uint8_t out[96], int nthreads) {
    orc_init();
    if (n == 0) { memset(out, 0, 96); return; }
    int c = pick_window(n);
    int W = (255 + c) / c + 1;          /* room for the carry of the signed recoding */
    if ((W - 1) * c >= 256) W--;      /* top digit cannot carry out: scalars are < 2^255 */
    u64 half = (u64)1 << (c - 1);
    g1a *P = (g1a *)malloc(sizeof(g1a) * n);
    int32_t *dig = (int32_t *)malloc(sizeof(int32_t) * n * (size_t)W);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        g1a_from_wire(&P[i], points + 96 * (u64)i);
        u64 k[FRN + 1]; scalar_canon(k, scalars + 32 * (u64)i); k[FRN] = 0;
        int carry = 0;
        for (int w = 0; w < W; w++) {
            int bit = w * c;
            u64 v = 0;
            if (bit < 256) {
                v = k[bit / 64] >> (bit % 64);
                if (bit % 64 + c > 64) v |= k[bit / 64 + 1] << (64 - bit % 64);
                v &= ((u64)1 << c) - 1;
            }
            int64_t d = (int64_t)v + carry;
            if ((u64)d > half) { d -= (int64_t)1 << c; carry = 1; } else carry = 0;
            dig[(size_t)i * W + w] = (int32_t)d;
        }
    }
    g1j *wsum = (g1j *)malloc(sizeof(g1j) * W);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int w = 0; w < W; w++) {
        g1j *bk = (g1j *)calloc(half, sizeof(g1j));
        for (u64 i = 0; i < n; i++) {
            int32_t d = dig[(size_t)i * W + w];
            if (d > 0) g1j_add_affine(&bk[d - 1], &bk[d - 1], &P[i], 0);
            else if (d < 0) g1j_add_affine(&bk[-d - 1], &bk[-d - 1], &P[i], 1);
        }
        g1j run, acc; g1j_set_inf(&run); g1j_set_inf(&acc);
        for (int64_t b = (int64_t)half - 1; b >= 0; b--) {
            g1j_add(&run, &run, &bk[b]);
            g1j_add(&acc, &acc, &run);
        }
        wsum[w] = acc;
        free(bk);
    }
    g1j tot; g1j_set_inf(&tot);
    for (int w = W - 1; w >= 0; w--) {
        for (int i = 0; i < c; i++) g1j_double(&tot, &tot);
        g1j_add(&tot, &tot, &wsum[w]);
    }
    g1j_to_wire(out, &tot);
    free(wsum); free(dig); free(P);
}

/* sum of n affine points (used to check the multi-GPU partial-sum combine) */
void orc_g1_sum(const uint8_t *points, u64 n, uint8_t out[96]) {
    orc_init();
    g1j acc; g1j_set_inf(&acc);
    for (u64 i = 0; i < n; i++) { g1a a; g1a_from_wire(&a, points + 96 * i); g1j_add_affine(&acc, &acc, &a, 0); }
    g1j_to_wire(out, &acc);
}

/* ------------------------------------------------------------------ synthetic inputs (shared definition with the GPU generator) */
static u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    u64 z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* base i = a_i * G with a_i = splitmix64(seed + i) (64-bit discrete log, known to the tests).
 * Reference definition, double-and-add: kept for the cross-check of the windowed version below. */
void orc_g1_synth_bases_naive(u64 seed, u64 start, u64 n, uint8_t *out, int nthreads) {
    orc_init();
    uint8_t gw[96]; orc_g1_generator(gw);
    g1a G; g1a_from_wire(&G, gw);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        u64 k[FRN] = {splitmix64(seed + start + (u64)i), 0, 0, 0};
        g1j j; g1j_mul_bits(&j, &G, k, 64);
        g1j_to_wire(out + 96 * (u64)i, &j);
    }
}
/* The same points, fast enough for 2^24 of them on the host (the CPU arm of bench.py synthesises its own bases): fixed-base
 * windows (8 x 255 affine multiples of G, one mixed addition per non-zero byte of a_i) and one shared inversion per block of
 * 256 points (Montgomery's trick) for the affine normalisation. */
void orc_g1_synth_bases(u64 seed, u64 start, u64 n, uint8_t *out, int nthreads) {
    orc_init();
    uint8_t gw[96]; orc_g1_generator(gw);
    g1a G; g1a_from_wire(&G, gw);
    static g1a table[8][256];
    static int table_done = 0;
#pragma omp critical(orc_synth_table)
    if (!table_done) {
        g1j base; g1j_from_affine(&base, &G);
        for (int w = 0; w < 8; w++) {
            g1j acc; g1j_set_inf(&acc);
            for (int d = 1; d < 256; d++) {
                g1j_add(&acc, &acc, &base);
                uint8_t wire[96]; g1j_to_wire(wire, &acc);
                g1a_from_wire(&table[w][d], wire);
            }
            for (int k = 0; k < 8; k++) g1j_double(&base, &base);
        }
        table_done = 1;
    }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
    enum { BLK = 256 };
    int64_t nblk = (int64_t)((n + BLK - 1) / BLK);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t b = 0; b < nblk; b++) {
        g1j pts[BLK]; fp pre[BLK];
        u64 i0 = (u64)b * BLK, cnt = n - i0 < BLK ? n - i0 : BLK;
        fp run; memcpy(run.l, FP_R, sizeof run.l);
        for (u64 t = 0; t < cnt; t++) {
            u64 k = splitmix64(seed + start + i0 + t);
            g1j acc; g1j_set_inf(&acc);
            for (int w = 0; w < 8; w++) {
                unsigned d = (unsigned)(k >> (8 * w)) & 255u;
                if (d) g1j_add_affine(&acc, &acc, &table[w][d], 0);
            }
            pts[t] = acc;
            pre[t] = run;                               /* product of the z of the earlier points of the block */
            if (!g1j_is_inf(&acc)) fp_mul(&run, &run, &acc.z);
        }
        fp inv; fp_inv(&inv, &run);
        for (int64_t t = (int64_t)cnt - 1; t >= 0; t--) {
            uint8_t *o = out + 96 * (i0 + (u64)t);
            if (g1j_is_inf(&pts[t])) { memset(o, 0, 96); continue; }
            fp zi, zi2, zi3, x, y;
            fp_mul(&zi, &inv, &pre[t]);                 /* 1 / z_t */
            fp_mul(&inv, &inv, &pts[t].z);
            fp_sqr(&zi2, &zi); fp_mul(&zi3, &zi2, &zi);
            fp_mul(&x, &pts[t].x, &zi2); fp_mul(&y, &pts[t].y, &zi3);
            fp_to_bytes(o, &x); fp_to_bytes(o + 48, &y);
        }
    }
}
/* scalar i: limbs w_j = splitmix64(seed + 4i + j), top limb masked to 63 bits (value < 2^255),
 * then one conditional subtraction of r */
void orc_fr_synth(u64 seed, u64 start, u64 n, uint8_t *out) {
    for (u64 i = 0; i < n; i++) {
        u64 k[FRN];
        for (int j = 0; j < FRN; j++) k[j] = splitmix64(seed + 4 * (start + i) + j);
        k[3] &= 0x7FFFFFFFFFFFFFFFULL;
        if (ge_n(k, FR_P, FRN)) sub_n(k, k, FR_P, FRN);
        memcpy(out + 32 * i, k, 32);
    }
}
/* sum_i s_i * a_i mod r, a_i 64-bit: the discrete log of MSM(s, a_i*G) */
void orc_fr_dot_u64(const uint8_t *scalars, const u64 *a, u64 n, uint8_t out[32]) {
    orc_init();
    fr acc; memset(&acc, 0, sizeof(acc));
    for (u64 i = 0; i < n; i++) {
        fr s, t; uint8_t ab[32] = {0};
        fr_from_bytes(&s, scalars + 32 * i);
        memcpy(ab, &a[i], 8);
        fr_from_bytes(&t, ab);
        fr_mul(&s, &s, &t);
        fr_add(&acc, &acc, &s);
    }
    fr_to_bytes(out, &acc);
}

/* ------------------------------------------------------------------ NTT over Fr */
/* O(n^2) DFT, X[k] = sum_i a[i] w^(ik): the definition halo2's best_fft computes */
void orc_ntt_naive(uint8_t *data, uint32_t log_n, const uint8_t omega[32]) {
    orc_init();
    u64 n = (u64)1 << log_n;
    fr *a = (fr *)malloc(sizeof(fr) * n), *o = (fr *)malloc(sizeof(fr) * n), w;
    for (u64 i = 0; i < n; i++) fr_from_bytes(&a[i], data + 32 * i);
    fr_from_bytes(&w, omega);
    fr wk; memcpy(wk.l, FR_R, 32);
    for (u64 k = 0; k < n; k++) {
        fr acc, x; memset(&acc, 0, sizeof(acc)); memcpy(x.l, FR_R, 32);
        for (u64 i = 0; i < n; i++) {
            fr t; fr_mul(&t, &a[i], &x); fr_add(&acc, &acc, &t); fr_mul(&x, &x, &wk);
        }
        o[k] = acc;
        fr_mul(&wk, &wk, &w);
    }
    for (u64 i = 0; i < n; i++) fr_to_bytes(data + 32 * i, &o[i]);
    free(a); free(o);
}

static u64 bitrev(u64 x, uint32_t bits) {
    u64 r = 0;
    for (uint32_t i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
/* in-place radix-2 decimation-in-time on Montgomery-form elements, natural order in/out */
static void ntt_core(fr *a, uint32_t log_n, const fr *w, int nthreads) {
    u64 n = (u64)1 << log_n;
    if (n == 1) return;
    for (u64 i = 0; i < n; i++) {
        u64 j = bitrev(i, log_n);
        if (i < j) { fr t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    fr *tw = (fr *)malloc(sizeof(fr) * (n / 2));
    memcpy(tw[0].l, FR_R, 32);
    for (u64 i = 1; i < n / 2; i++) fr_mul(&tw[i], &tw[i - 1], w);
    for (uint32_t s = 0; s < log_n; s++) {
        u64 m = (u64)1 << s, step = n >> (s + 1);
#pragma omp parallel for num_threads(nthreads) schedule(static) if (n >= 4096)
        for (int64_t idx = 0; idx < (int64_t)(n / 2); idx++) {
            u64 j = (u64)idx & (m - 1), blk = (u64)idx >> s;
            u64 lo = blk * 2 * m + j, hi = lo + m;
            fr t; fr_mul(&t, &a[hi], &tw[j * step]);
            fr u = a[lo];
            fr_add(&a[lo], &u, &t);
            fr_sub(&a[hi], &u, &t);
        }
    }
    free(tw);
}
/* flags: bit0 = scale the result by 1/n (caller passes omega^-1 for an inverse transform);
 * coset_in  != NULL: a[i] *= g^i before the transform (coeff_to_extended);
 * coset_out != NULL: a[i] *= g^i after it (extended_to_coeff passes g^-1). */
void orc_ntt(uint8_t *data, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
             const uint8_t *coset_in, const uint8_t *coset_out, int nthreads) {
    orc_init();
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    u64 n = (u64)1 << log_n;
    fr *a = (fr *)malloc(sizeof(fr) * n), w;
#pragma omp parallel for num_threads(nthreads) schedule(static) if (n >= 4096)
    for (int64_t i = 0; i < (int64_t)n; i++) fr_from_bytes(&a[i], data + 32 * (u64)i);
    fr_from_bytes(&w, omega);
    if (coset_in) {
        fr g, x; fr_from_bytes(&g, coset_in); memcpy(x.l, FR_R, 32);
        for (u64 i = 0; i < n; i++) { fr_mul(&a[i], &a[i], &x); fr_mul(&x, &x, &g); }
    }
    ntt_core(a, log_n, &w, nthreads);
    if (flags & 1) {
        uint8_t nb[32] = {0}; u64 nn = n; memcpy(nb, &nn, 8);
        fr ninv; fr_from_bytes(&ninv, nb); fr_inv(&ninv, &ninv);
        for (u64 i = 0; i < n; i++) fr_mul(&a[i], &a[i], &ninv);
    }
    if (coset_out) {
        fr g, x; fr_from_bytes(&g, coset_out); memcpy(x.l, FR_R, 32);
        for (u64 i = 0; i < n; i++) { fr_mul(&a[i], &a[i], &x); fr_mul(&x, &x, &g); }
    }
#pragma omp parallel for num_threads(nthreads) schedule(static) if (n >= 4096)
    for (int64_t i = 0; i < (int64_t)n; i++) fr_to_bytes(data + 32 * (u64)i, &a[i]);
    free(a);
}

/* ------------------------------------------------------------------ polynomial side
 * Plain serial restatements, the checker for the Fr vector kernels at sizes the big-int oracle cannot walk.
 * All vectors are canonical 32-byte little-endian elements. */
static int fr_is_zero_(const fr *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }

/* synthetic division by (X - z): quot[i-1] = s_i, eval = s_0, s_i = c_i + z s_{i+1}  (the witness polynomials of the
 * KZG multi-open, /root/reference/src/plutus_gen/extraction/pcs/kzg.rs:55-79) */
void orc_fr_kate_div(const uint8_t *coeffs, u64 n, const uint8_t z[32], uint8_t *quot, uint8_t eval[32]) {
    orc_init();
    fr zz, s, c;
    fr_from_bytes(&zz, z);
    memset(&s, 0, sizeof(s));
    for (u64 k = n; k-- > 0;) {
        fr_from_bytes(&c, coeffs + 32 * k);
        fr_mul(&s, &s, &zz);
        fr_add(&s, &s, &c);
        if (k && quot) fr_to_bytes(quot + 32 * (k - 1), &s);
    }
    if (eval) fr_to_bytes(eval, &s);
}
/* out[i] = init * prod_{j<i} in[j] (exclusive) or prod_{j<=i} (inclusive): the grand products of the permutation /
 * lookup arguments (/root/reference/src/plutus_gen/extraction/data/extraction_steps/permutation.rs) */
void orc_fr_running_product(const uint8_t *in, u64 n, const uint8_t *init, int inclusive, uint8_t *out) {
    orc_init();
    fr acc, v, nxt;
    if (init) fr_from_bytes(&acc, init);
    else memcpy(acc.l, FR_R, 32);
    for (u64 i = 0; i < n; i++) {
        fr_from_bytes(&v, in + 32 * i);
        fr_mul(&nxt, &acc, &v);
        fr_to_bytes(out + 32 * i, inclusive ? &nxt : &acc);
        acc = nxt;
    }
}
/* elementwise inverse, zeros stay zero (halo2 batch_invert); Montgomery's trick over the whole vector */
void orc_fr_batch_invert(const uint8_t *in, u64 n, uint8_t *out) {
    orc_init();
    fr *pre = (fr *)malloc(sizeof(fr) * (n ? n : 1)), run, v, inv;
    memcpy(run.l, FR_R, 32);
    for (u64 i = 0; i < n; i++) {
        pre[i] = run;
        fr_from_bytes(&v, in + 32 * i);
        if (!fr_is_zero_(&v)) fr_mul(&run, &run, &v);
    }
    fr_inv(&inv, &run);
    for (u64 i = n; i-- > 0;) {
        fr_from_bytes(&v, in + 32 * i);
        if (fr_is_zero_(&v)) { memset(out + 32 * i, 0, 32); continue; }
        fr r;
        fr_mul(&r, &inv, &pre[i]);
        fr_to_bytes(out + 32 * i, &r);
        fr_mul(&inv, &inv, &v);
    }
    free(pre);
}
/* out = sum_k coeffs[k] * polys[k] (polys stored back to back, n elements each) */
void orc_fr_lincomb(const uint8_t *polys, const uint8_t *coeffs, u64 count, u64 n, uint8_t *out) {
    orc_init();
    for (u64 i = 0; i < n; i++) {
        fr acc, c, v;
        memset(&acc, 0, sizeof(acc));
        for (u64 k = 0; k < count; k++) {
            fr_from_bytes(&c, coeffs + 32 * k);
            fr_from_bytes(&v, polys + 32 * (k * n + i));
            fr_mul(&v, &v, &c);
            fr_add(&acc, &acc, &v);
        }
        fr_to_bytes(out + 32 * i, &acc);
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
