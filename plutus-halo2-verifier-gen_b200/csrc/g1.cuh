// g1.cuh -- BLS12-381 G1 (y^2 = x^3 + 4 over Fp) point arithmetic for the MSM kernels.
//
// Replaces the blst G1 add / double / mixed-add routines that midnight-curves' G1Projective
// reaches (call sites: every commit/commit_lagrange behind create_proof,
// /root/reference/examples/simple_mul.rs:72; the verifier fold `eval`,
// /root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/MSMEval.hs:20-25).
//
// Bucket accumulators use extended Jacobian ("XYZZ") coordinates: x = X/ZZ, y = Y/ZZZ with
// ZZ^3 = ZZZ^2, ZZ = 0 for the identity.  A mixed addition of an affine base costs
// 8M + 2S (EFD madd-2008-s), the cheapest complete-enough formula when one operand is affine.
// Affine points use the reference's convention for the identity: (0, 0)
// (/root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/CompressUncompress.hs:72).
#pragma once
#include "field.cuh"

namespace b200zk {

struct G1Affine {
    Fp x, y;
};
struct G1Xyzz {
    Fp x, y, zz, zzz;
};

// how the point formulas obtain their field products: inlined (hot kernels tune this) or, in
// msm.cuh, an out-of-line call that keeps the code small
struct MulInline {
    static HD Fp mul(const Fp& a, const Fp& b) { return fe_mul(a, b); }
    static HD Fp sqr(const Fp& a) { return fe_mul(a, a); }
    static HD Fp mulsub(const Fp& a, const Fp& b, const Fp& c, const Fp& d) { return fe_sub(fe_mul(a, b), fe_mul(c, d)); }
};

HD bool g1a_is_inf(const G1Affine& a) { return fe_is_zero(a.x) && fe_is_zero(a.y); }
HD bool xyzz_is_inf(const G1Xyzz& a) { return fe_is_zero(a.zz); }
HD void xyzz_set_inf(G1Xyzz& a) {
    a.x = fe_zero<FpParams>();
    a.y = fe_zero<FpParams>();
    a.zz = fe_zero<FpParams>();
    a.zzz = fe_zero<FpParams>();
}
HD void xyzz_from_affine(G1Xyzz& r, const G1Affine& a, bool neg) {
    if (g1a_is_inf(a)) { xyzz_set_inf(r); return; }
    r.x = a.x;
    r.y = neg ? fe_neg(a.y) : a.y;
    r.zz = fe_one<FpParams>();
    r.zzz = fe_one<FpParams>();
}

// acc = 2 * (affine q)   (EFD mdbl-2008-s-1), q not the identity, y != 0 (no 2-torsion in G1)
HD void xyzz_dbl_affine(G1Xyzz& r, const Fp& qx, const Fp& qy) {
    Fp u = fe_dbl(qy);
    Fp v = fe_sqr(u);
    Fp w = fe_mul(u, v);
    Fp s = fe_mul(qx, v);
    Fp m = fe_sqr(qx);
    m = fe_add(fe_dbl(m), m);
    Fp x3 = fe_sub(fe_sub(fe_sqr(m), s), s);
    Fp y3 = fe_sub(fe_mul(m, fe_sub(s, x3)), fe_mul(w, qy));
    r.x = x3; r.y = y3; r.zz = v; r.zzz = w;
}

// acc = 2*acc (EFD dbl-2008-s-1, a = 0)
template <class M> HD void xyzz_dbl_t(G1Xyzz& a) {
    if (xyzz_is_inf(a)) return;
    Fp u = fe_dbl(a.y);
    Fp v = M::mul(u, u);
    Fp w = M::mul(u, v);
    Fp s = M::mul(a.x, v);
    Fp m = M::mul(a.x, a.x);
    m = fe_add(fe_dbl(m), m);
    Fp x3 = fe_sub(fe_sub(M::mul(m, m), s), s);
    Fp y3 = fe_sub(M::mul(m, fe_sub(s, x3)), M::mul(w, a.y));
    a.zz = M::mul(v, a.zz);
    a.zzz = M::mul(w, a.zzz);
    a.x = x3; a.y = y3;
}
HD void xyzz_dbl(G1Xyzz& a) { xyzz_dbl_t<MulInline>(a); }

// acc += (neg ? -q : q), q affine (EFD madd-2008-s: 8M + 2S)
HD void xyzz_add_mixed(G1Xyzz& acc, const G1Affine& q, bool neg) {
    if (g1a_is_inf(q)) return;
    Fp qy = neg ? fe_neg(q.y) : q.y;
    if (xyzz_is_inf(acc)) {
        acc.x = q.x; acc.y = qy; acc.zz = fe_one<FpParams>(); acc.zzz = fe_one<FpParams>();
        return;
    }
    Fp u2 = fe_mul(q.x, acc.zz);
    Fp s2 = fe_mul(qy, acc.zzz);
    Fp p = fe_sub(u2, acc.x);
    Fp r = fe_sub(s2, acc.y);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) xyzz_dbl_affine(acc, q.x, qy);   // same point
        else xyzz_set_inf(acc);                              // opposite points
        return;
    }
    Fp pp = fe_sqr(p);
    Fp ppp = fe_mul(p, pp);
    Fp qq = fe_mul(acc.x, pp);
    Fp x3 = fe_sub(fe_sub(fe_sub(fe_sqr(r), ppp), qq), qq);
    Fp y3 = fe_sub(fe_mul(r, fe_sub(qq, x3)), fe_mul(acc.y, ppp));
    acc.zz = fe_mul(acc.zz, pp);
    acc.zzz = fe_mul(acc.zzz, ppp);
    acc.x = x3; acc.y = y3;
}

// acc += q for an affine q whose sign is already applied; products through the policy M
// (same formulas as xyzz_add_mixed, ordered to keep few values live)
template <class M> HD void xyzz_add_mixed_pol(G1Xyzz& acc, const G1Affine& q) {
    if (g1a_is_inf(q)) return;
    if (xyzz_is_inf(acc)) {
        acc.x = q.x; acc.y = q.y; acc.zz = fe_one<FpParams>(); acc.zzz = fe_one<FpParams>();
        return;
    }
    Fp p = fe_sub(M::mul(q.x, acc.zz), acc.x);
    Fp r = fe_sub(M::mul(q.y, acc.zzz), acc.y);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) {
            G1Xyzz d;
            d.x = q.x; d.y = q.y; d.zz = fe_one<FpParams>(); d.zzz = fe_one<FpParams>();
            xyzz_dbl_t<M>(d);
            acc = d;
        } else xyzz_set_inf(acc);
        return;
    }
    Fp pp = M::mul(p, p);
    Fp qq = M::mul(acc.x, pp);
    acc.zz = M::mul(acc.zz, pp);
    Fp ppp = M::mul(p, pp);
    acc.zzz = M::mul(acc.zzz, ppp);
    Fp t = M::mul(acc.y, ppp);
    Fp x3 = fe_sub(fe_sub(fe_sub(M::mul(r, r), ppp), qq), qq);
    acc.y = fe_sub(M::mul(r, fe_sub(qq, x3)), t);
    acc.x = x3;
}

// acc += q, both XYZZ (EFD add-2008-s: 12M + 2S)
template <class M> HD void xyzz_add_t(G1Xyzz& acc, const G1Xyzz& q) {
    if (xyzz_is_inf(q)) return;
    if (xyzz_is_inf(acc)) { acc = q; return; }
    Fp u1 = M::mul(acc.x, q.zz);
    Fp u2 = M::mul(q.x, acc.zz);
    Fp s1 = M::mul(acc.y, q.zzz);
    Fp s2 = M::mul(q.y, acc.zzz);
    Fp p = fe_sub(u2, u1);
    Fp r = fe_sub(s2, s1);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) xyzz_dbl_t<M>(acc);
        else xyzz_set_inf(acc);
        return;
    }
    Fp pp = M::mul(p, p);
    Fp ppp = M::mul(p, pp);
    Fp qq = M::mul(u1, pp);
    Fp x3 = fe_sub(fe_sub(fe_sub(M::mul(r, r), ppp), qq), qq);
    Fp y3 = fe_sub(M::mul(r, fe_sub(qq, x3)), M::mul(s1, ppp));
    acc.zz = M::mul(M::mul(acc.zz, q.zz), pp);
    acc.zzz = M::mul(M::mul(acc.zzz, q.zzz), ppp);
    acc.x = x3; acc.y = y3;
}
HD void xyzz_add(G1Xyzz& acc, const G1Xyzz& q) { xyzz_add_t<MulInline>(acc, q); }

// affine normalisation: one inversion for both denominators
HD G1Affine xyzz_to_affine(const G1Xyzz& a) {
    G1Affine r;
    if (xyzz_is_inf(a)) { r.x = fe_zero<FpParams>(); r.y = fe_zero<FpParams>(); return r; }
    Fp t = fe_inv_gcd(fe_mul(a.zz, a.zzz));
    Fp izz = fe_mul(t, a.zzz);
    Fp izzz = fe_mul(t, a.zz);
    r.x = fe_mul(a.x, izz);
    r.y = fe_mul(a.y, izzz);
    return r;
}

// y^2 == x^3 + 4 (Montgomery form), identity accepted
HD bool g1a_on_curve(const G1Affine& a) {
    if (g1a_is_inf(a)) return true;
    Fp four = fe_dbl(fe_dbl(fe_one<FpParams>()));
    return fe_eq(fe_sqr(a.y), fe_add(fe_mul(fe_sqr(a.x), a.x), four));
}

}  // namespace b200zk
