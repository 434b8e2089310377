// CPU-side check of the device field / curve algorithms (field.cuh, g1.cuh compiled for the
// host with the carry flag emulated) against the oracle.  Test infrastructure only.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <random>
#include "field.cuh"
#include "g1.cuh"
extern "C" {
void orc_init(void);
void orc_fp_mul(const uint8_t*, const uint8_t*, uint8_t*);
void orc_fp_add(const uint8_t*, const uint8_t*, uint8_t*);
void orc_fp_sub(const uint8_t*, const uint8_t*, uint8_t*);
void orc_fp_inv(const uint8_t*, uint8_t*);
void orc_fr_mul(const uint8_t*, const uint8_t*, uint8_t*);
void orc_fr_add(const uint8_t*, const uint8_t*, uint8_t*);
void orc_fr_sub(const uint8_t*, const uint8_t*, uint8_t*);
void orc_fr_inv(const uint8_t*, uint8_t*);
void orc_fp_to_mont(const uint8_t*, uint8_t*);
void orc_fr_to_mont(const uint8_t*, uint8_t*);
void orc_g1_generator(uint8_t*);
void orc_g1_mul(const uint8_t*, const uint8_t*, uint8_t*);
void orc_g1_add(const uint8_t*, const uint8_t*, uint8_t*);
}
using namespace b200zk;
static std::mt19937_64 rng(12345);
template <class P> static Fe<P> rnd_canon() {
    Fe<P> a;
    for (int i = 0; i < P::N; i++) a.l[i] = (uint32_t)rng();
    a.l[P::N - 1] &= (P::N == 12) ? 0x0fffffffu : 0x3fffffffu;  // < modulus for both fields
    return a;
}
template <class P> static int check_field(const char* name,
        void (*omul)(const uint8_t*, const uint8_t*, uint8_t*), void (*oadd)(const uint8_t*, const uint8_t*, uint8_t*),
        void (*osub)(const uint8_t*, const uint8_t*, uint8_t*), void (*oinv)(const uint8_t*, uint8_t*),
        void (*omont)(const uint8_t*, uint8_t*)) {
    int bad = 0;
    for (int it = 0; it < 20000; it++) {
        Fe<P> a = rnd_canon<P>(), b = rnd_canon<P>();
        if (it == 0) { a = fe_zero<P>(); }
        if (it == 1) { for (int i = 0; i < P::N; i++) a.l[i] = P::p(i); a.l[0] -= 1; b = a; }   // p-1
        if (it == 2) { a = fe_zero<P>(); a.l[0] = 1; }
        Fe<P> am = fe_to_mont(a), bm = fe_to_mont(b), chk;
        omont((uint8_t*)a.l, (uint8_t*)chk.l);
        if (!fe_eq(am, chk)) { bad++; printf("%s to_mont mismatch it=%d\n", name, it); }
        Fe<P> r = fe_from_mont(fe_mul(am, bm)), e;
        omul((uint8_t*)a.l, (uint8_t*)b.l, (uint8_t*)e.l);
        if (!fe_eq(r, e)) { bad++; if (bad < 5) printf("%s mul mismatch it=%d\n", name, it); }
        r = fe_from_mont(fe_sqr_fast(am)); omul((uint8_t*)a.l, (uint8_t*)a.l, (uint8_t*)e.l);
        if (!fe_eq(r, e) || !fe_eq(fe_sqr_fast(am), fe_mul(am, am))) { bad++; if (bad < 5) printf("%s sqr_fast mismatch it=%d\n", name, it); }
        if constexpr (P::N == 12) {   // a*b + c*d in one pass (Fp only: needs three bits of headroom), against the two-product formulation (also with both products at their maximum)
            Fe<P> cm = fe_to_mont(rnd_canon<P>()), dm = fe_to_mont(rnd_canon<P>());
            if (it == 1) { cm = am; dm = am; }
            Fe<P> want = fe_add(fe_mul(am, bm), fe_mul(cm, dm));
            if (!fe_eq(fe_mul2(am, bm, cm, dm), want)) { bad++; if (bad < 5) printf("%s mul2 mismatch it=%d\n", name, it); }
            // raw maximal operands (all limbs p-1 pattern, not Montgomery-converted): stresses the accumulator bound
            Fe<P> mx; for (int i = 0; i < P::N; i++) mx.l[i] = P::p(i); mx.l[0] -= 1;
            if (it < 4 && !fe_eq(fe_mul2(mx, mx, mx, mx), fe_add(fe_mul(mx, mx), fe_mul(mx, mx)))) { bad++; printf("%s mul2 max mismatch\n", name); }
        }
        r = fe_from_mont(fe_add(am, bm)); oadd((uint8_t*)a.l, (uint8_t*)b.l, (uint8_t*)e.l);
        if (!fe_eq(r, e)) { bad++; if (bad < 5) printf("%s add mismatch it=%d\n", name, it); }
        r = fe_from_mont(fe_sub(am, bm)); osub((uint8_t*)a.l, (uint8_t*)b.l, (uint8_t*)e.l);
        if (!fe_eq(r, e)) { bad++; if (bad < 5) printf("%s sub mismatch it=%d\n", name, it); }
        {   // the NTT butterfly's product of a twiddle with an uncorrected difference: w * (x - y + p), difference in [0, 2p)
            Fe<P> w = fe_to_mont(rnd_canon<P>()), x = am, y = bm;
            Fe<P> pm1; for (int i = 0; i < P::N; i++) pm1.l[i] = P::p(i); pm1.l[0] -= 1;
            if (it == 3) { x = pm1; y = fe_zero<P>(); w = pm1; }      // largest difference (2p - 1), largest twiddle
            if (it == 4) { x = fe_zero<P>(); y = pm1; }               // smallest
            if (it == 5) { y = x; }                                    // zero difference, represented as p
            if (!fe_eq(fe_mul(w, fe_sub_lazy(x, y)), fe_mul(fe_sub(x, y), w))) { bad++; if (bad < 5) printf("%s sub_lazy mismatch it=%d\n", name, it); }
        }
        if (it < 50 && !fe_is_zero(a)) {
            r = fe_from_mont(fe_inv(am)); oinv((uint8_t*)a.l, (uint8_t*)e.l);
            if (!fe_eq(r, e)) { bad++; printf("%s inv mismatch it=%d\n", name, it); }
        }
        if (it < 3000 && !fe_is_zero(a)) {
            r = fe_from_mont(fe_inv_gcd(am)); oinv((uint8_t*)a.l, (uint8_t*)e.l);
            if (!fe_eq(r, e)) { bad++; if (bad < 5) printf("%s inv_gcd mismatch it=%d\n", name, it); }
            r = fe_from_mont(fe_inv_fast(am));
            if (!fe_eq(r, e)) { bad++; if (bad < 5) printf("%s inv_fast mismatch it=%d\n", name, it); }
        }
    }
    printf("%s: %s\n", name, bad ? "FAIL" : "ok");
    return bad;
}
static G1Affine from_wire(const uint8_t* w) {
    G1Affine a; Fp t;
    memcpy(t.l, w, 48); a.x = fe_to_mont(t);
    memcpy(t.l, w + 48, 48); a.y = fe_to_mont(t);
    return a;
}
static void to_wire(uint8_t* w, const G1Affine& a) {
    Fp t = fe_from_mont(a.x); memcpy(w, t.l, 48);
    t = fe_from_mont(a.y); memcpy(w + 48, t.l, 48);
}
int main() {
    orc_init();
    int bad = 0;
    bad += check_field<FpParams>("Fp", orc_fp_mul, orc_fp_add, orc_fp_sub, orc_fp_inv, orc_fp_to_mont);
    bad += check_field<FrParams>("Fr", orc_fr_mul, orc_fr_add, orc_fr_sub, orc_fr_inv, orc_fr_to_mont);
    // curve: k*G via double-and-add in XYZZ (mixed adds + doublings), full adds, special cases
    uint8_t gw[96], ew[96], rw[96];
    orc_g1_generator(gw);
    G1Affine G = from_wire(gw);
    if (!g1a_on_curve(G)) { bad++; printf("generator not on curve\n"); }
    for (int it = 0; it < 40; it++) {
        uint8_t s[32] = {0};
        uint64_t k = rng() >> (it % 40);
        if (it == 0) k = 1; if (it == 1) k = 2; if (it == 2) k = 3;
        memcpy(s, &k, 8);
        G1Xyzz acc; xyzz_set_inf(acc);
        for (int b = 63; b >= 0; b--) { xyzz_dbl(acc); if ((k >> b) & 1) xyzz_add_mixed(acc, G, false); }
        to_wire(rw, xyzz_to_affine(acc));
        orc_g1_mul(gw, s, ew);
        if (memcmp(rw, ew, 96)) { bad++; printf("g1 mul mismatch k=%llu\n", (unsigned long long)k); }
        // full add: acc + acc2 where acc2 = 5G (xyzz), compare with oracle add
        G1Xyzz five; xyzz_set_inf(five);
        for (int j = 0; j < 5; j++) xyzz_add_mixed(five, G, false);   // exercises inf->set, dbl (2nd add), generic
        uint8_t fw[96], s5[32] = {5}; orc_g1_mul(gw, s5, fw);
        G1Xyzz sum = acc; xyzz_add(sum, five);
        to_wire(rw, xyzz_to_affine(sum)); orc_g1_add(ew, fw, ew);
        if (memcmp(rw, ew, 96)) { bad++; printf("g1 add mismatch k=%llu\n", (unsigned long long)k); }
        // P + (-P) = inf (mixed and full), P + P via full add
        G1Xyzz t = acc; G1Affine pa = xyzz_to_affine(acc);
        xyzz_add_mixed(t, pa, true);
        if (!xyzz_is_inf(t)) { bad++; printf("P-P mixed not inf\n"); }
        G1Xyzz n; xyzz_from_affine(n, pa, true); t = acc; xyzz_add(t, n);
        if (!xyzz_is_inf(t)) { bad++; printf("P-P full not inf\n"); }
        t = acc; xyzz_add(t, acc); G1Xyzz d = acc; xyzz_dbl(d);
        uint8_t a1[96], a2[96]; to_wire(a1, xyzz_to_affine(t)); to_wire(a2, xyzz_to_affine(d));
        if (memcmp(a1, a2, 96)) { bad++; printf("P+P full != dbl\n"); }
        t = acc; xyzz_add_mixed(t, pa, false); to_wire(a1, xyzz_to_affine(t));
        if (memcmp(a1, a2, 96)) { bad++; printf("P+P mixed != dbl\n"); }
    }
    printf("G1: %s\n", bad ? "FAIL" : "ok");
    return bad ? 1 : 0;
}
