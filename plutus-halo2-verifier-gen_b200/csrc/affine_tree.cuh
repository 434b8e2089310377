// affine_tree.cuh -- bucket accumulation by batched affine additions.
//
// A mixed XYZZ addition costs 10 field products per point.  An affine addition costs 3 once the
// inverse of (x2 - x1) is known, and Montgomery's trick shares one inversion among all the
// independent additions of a batch at 3 more products each.  The points of one task (a run of
// entries of one bucket) are therefore summed as a pairwise tree that is local to the thread:
//   round r:  p[2j] + p[2j+1] -> q[j]  for all j, one shared inversion (fe_inv_fast, branch-free,
//             so the 32 lanes of a warp -- which hold tasks of equal length -- stay in lockstep);
//             an odd last point is folded into the thread's XYZZ accumulator;
//   the rounds stop when fewer than AFF_MIN_PAIRS pairs remain (the inversion would no longer pay)
//   and the survivors are folded into the XYZZ accumulator with ordinary mixed additions.
// Per point that is ~6.2 products in the large rounds instead of 10.
//
// Intermediate points live in two ping-pong scratch arrays addressed by (start >> 1) + j, which
// keeps the slices of different tasks disjoint at every round; the prefix products of a round use
// a third array with the same addressing.  Equal points (doubling), opposite points and identity
// points inside a pair are handled explicitly -- the adversarial inputs of the parity tests
// (repeated bases, P and -P, (0,0)) all come through here.
//
// Everything is __host__ __device__ so that tests/host/host_affine_tree_test.cpp runs the same code
// against the oracle on a GPU-less box.
#pragma once
#include "g1.cuh"

namespace b200zk {

constexpr uint32_t AFF_MIN_PAIRS = 12;   // a round needs at least this many pairs to beat mixed additions

HD Fp aff_ld(const uint32_t* p) {
    Fp r;
#ifdef __CUDA_ARCH__
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
#else
    for (int i = 0; i < 12; i++) r.l[i] = p[i];
#endif
    return r;
}
HD void aff_st(uint32_t* p, const Fp& v) {
#ifdef __CUDA_ARCH__
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
#else
    for (int i = 0; i < 12; i++) p[i] = v.l[i];
#endif
}

struct AffTreeMem {
    const uint32_t* bases;     // packed Montgomery affine rows (24 limbs per point)
    const uint32_t* entries;   // point index | sign << 31, sorted by bucket
    uint32_t* scr[2];          // ping-pong intermediate points, 24 limbs per slot, >= M/2 + 1 slots each
    uint32_t* pre;             // prefix products of a round, 12 limbs per slot, >= M/2 + 1 slots
};

// point i of the current round: from the base table through the entry list (round 0) or from scratch
HD G1Affine aff_point(const AffTreeMem& m, int src, uint32_t start, uint32_t sbase, uint32_t i) {
    G1Affine p;
    if (src < 0) {
        uint32_t e = m.entries[start + i];
        const uint32_t* b = m.bases + 24 * (uint64_t)(e & 0x7fffffffu);
        p.x = aff_ld(b);
        p.y = aff_ld(b + 12);
        if ((e >> 31) && !fe_is_zero(p.y)) p.y = fe_neg(p.y);
    } else {
        const uint32_t* b = m.scr[src] + 24 * (uint64_t)(sbase + i);
        p.x = aff_ld(b);
        p.y = aff_ld(b + 12);
    }
    return p;
}

// kind of a pair and the denominator its slope needs (1 when no slope is needed)
enum { AFF_ADD = 0, AFF_DBL = 1, AFF_INF = 2, AFF_COPY_A = 3, AFF_COPY_B = 4 };
HD int aff_classify(const G1Affine& a, const G1Affine& b, Fp& den) {
    bool ai = g1a_is_inf(a), bi = g1a_is_inf(b);
    den = fe_one<FpParams>();
    if (ai && bi) return AFF_INF;
    if (ai) return AFF_COPY_B;
    if (bi) return AFF_COPY_A;
    if (fe_eq(a.x, b.x)) {
        if (fe_eq(a.y, b.y)) { den = fe_dbl(a.y); return AFF_DBL; }   // y != 0: E(Fp) has odd order
        return AFF_INF;                                                // opposite points
    }
    den = fe_sub(b.x, a.x);
    return AFF_ADD;
}

// Sums the `len` entries [start, start+len) of one bucket into acc (XYZZ).  M supplies the products.
template <class M, class INV>
HD void msm_affine_tree_task(const AffTreeMem& mem, uint32_t start, uint32_t len, G1Xyzz& acc) {
    xyzz_set_inf(acc);
    const uint32_t sbase = start >> 1;
    uint32_t m = len;
    int src = -1;                       // -1: base table via entries; 0/1: scratch array
    while ((m >> 1) >= AFF_MIN_PAIRS) {
        const uint32_t np = m >> 1;
        const int dst = src < 0 ? 0 : (src ^ 1);
        if (m & 1) {                    // odd one out joins the accumulator directly
            G1Affine last = aff_point(mem, src, start, sbase, m - 1);
            xyzz_add_mixed_pol<M>(acc, last);
        }
        // pass 1: running product of the denominators
        Fp run = fe_one<FpParams>();
        for (uint32_t j = 0; j < np; j++) {
            G1Affine a = aff_point(mem, src, start, sbase, 2 * j), b = aff_point(mem, src, start, sbase, 2 * j + 1);
            Fp den;
            aff_classify(a, b, den);
            aff_st(mem.pre + 12 * (uint64_t)(sbase + j), run);
            run = M::mul(run, den);
        }
        Fp inv = INV::inv(run);
        // pass 2: peel the inverses off from the back and finish the additions
        for (uint32_t j = np; j-- > 0;) {
            G1Affine a = aff_point(mem, src, start, sbase, 2 * j), b = aff_point(mem, src, start, sbase, 2 * j + 1);
            Fp den;
            int kind = aff_classify(a, b, den);
            Fp dinv = M::mul(inv, aff_ld(mem.pre + 12 * (uint64_t)(sbase + j)));
            inv = M::mul(inv, den);
            G1Affine r;
            if (kind == AFF_ADD || kind == AFF_DBL) {
                Fp num;
                if (kind == AFF_ADD) num = fe_sub(b.y, a.y);
                else { Fp xx = M::mul(a.x, a.x); num = fe_add(fe_dbl(xx), xx); }
                Fp lam = M::mul(num, dinv);
                Fp x3 = fe_sub(fe_sub(M::mul(lam, lam), a.x), b.x);
                r.x = x3;
                r.y = fe_sub(M::mul(lam, fe_sub(a.x, x3)), a.y);
            } else if (kind == AFF_COPY_A) r = a;
            else if (kind == AFF_COPY_B) r = b;
            else { r.x = fe_zero<FpParams>(); r.y = fe_zero<FpParams>(); }
            uint32_t* o = mem.scr[dst] + 24 * (uint64_t)(sbase + j);
            aff_st(o, r.x);
            aff_st(o + 12, r.y);
        }
        m = np;
        src = dst;
    }
    for (uint32_t i = 0; i < m; i++) {
        G1Affine p = aff_point(mem, src, start, sbase, i);
        xyzz_add_mixed_pol<M>(acc, p);
    }
}

}  // namespace b200zk
