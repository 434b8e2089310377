// CPU-side check of the batched-affine bucket accumulation (affine_tree.cuh compiled for the host)
// against the oracle: random runs, repeated points, opposite points, identity points, all lengths
// around the round thresholds.  Test infrastructure only.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <random>
#include <vector>
#include "affine_tree.cuh"
extern "C" {
void orc_init(void);
void orc_g1_synth_bases(uint64_t seed, uint64_t start, uint64_t n, uint8_t* out, int nthreads);
void orc_g1_msm(const uint8_t* points, const uint8_t* scalars, uint64_t n, uint8_t* out, int nthreads);
}
using namespace b200zk;
struct InvHost { static Fp inv(const Fp& a) { return fe_inv_fast(a); } };
static const uint8_t R_MINUS_1[32] = {0x00,0x00,0x00,0x00,0xff,0xff,0xff,0xff,0xfe,0x5b,0xfe,0xff,0x02,0xa4,0xbd,0x53,
                                      0x05,0xd8,0xa1,0x09,0x08,0xd8,0x39,0x33,0x48,0x7d,0x9d,0x29,0x53,0xa7,0xed,0x73};
int main() {
    orc_init();
    std::mt19937_64 rng(7);
    const int NB = 64;                       // distinct base points, index NB = identity
    std::vector<uint8_t> wire(96 * (NB + 1), 0);
    orc_g1_synth_bases(0xB200, 0, NB, wire.data(), 1);
    std::vector<uint32_t> bases(24 * (NB + 1), 0);
    for (int i = 0; i < NB; i++) {
        Fp x, y;
        memcpy(x.l, &wire[96 * i], 48); memcpy(y.l, &wire[96 * i + 48], 48);
        x = fe_to_mont(x); y = fe_to_mont(y);
        memcpy(&bases[24 * i], x.l, 48); memcpy(&bases[24 * i + 12], y.l, 48);
    }
    int bad = 0;
    for (int it = 0; it < 400; it++) {
        uint32_t start = (uint32_t)(rng() % 50);
        uint32_t len = 1 + (uint32_t)(rng() % (it < 100 ? 60 : 400));
        if (it < 8) len = 22 + it;           // around 2 * AFF_MIN_PAIRS
        std::vector<uint32_t> entries(start + len);
        int mode = it % 5;                   // 0: random  1: few distinct (many repeats)  2: P,-P pairs  3: with identities  4: all the same point
        for (uint32_t i = 0; i < start + len; i++) {
            uint32_t idx = (uint32_t)(rng() % NB), neg = (uint32_t)(rng() & 1);
            if (mode == 1) idx %= 3;
            if (mode == 2 && (i & 1)) { idx = entries[i - 1] & 0x7fffffffu; neg = (entries[i - 1] >> 31) ^ 1; }
            if (mode == 3 && (rng() % 4 == 0)) idx = NB;
            if (mode == 4) { idx = 5; neg = 0; }
            entries[i] = idx | (neg << 31);
        }
        size_t slots = (start + len) / 2 + 2;
        std::vector<uint32_t> sA(24 * slots), sB(24 * slots), pre(12 * slots);
        AffTreeMem mem{bases.data(), entries.data(), {sA.data(), sB.data()}, pre.data()};
        G1Xyzz acc;
        msm_affine_tree_task<MulInline, InvHost>(mem, start, len, acc);
        G1Affine got = xyzz_to_affine(acc);
        uint8_t gw[96];
        Fp t = fe_from_mont(got.x); memcpy(gw, t.l, 48);
        t = fe_from_mont(got.y); memcpy(gw + 48, t.l, 48);
        // oracle: MSM with scalars +1 / -1
        std::vector<uint8_t> pts(96 * len), sc(32 * len, 0);
        for (uint32_t i = 0; i < len; i++) {
            uint32_t e = entries[start + i];
            memcpy(&pts[96 * i], &wire[96 * (e & 0x7fffffffu)], 96);
            if (e >> 31) memcpy(&sc[32 * i], R_MINUS_1, 32); else sc[32 * i] = 1;
        }
        uint8_t ew[96];
        orc_g1_msm(pts.data(), sc.data(), len, ew, 1);
        if (memcmp(gw, ew, 96)) { bad++; if (bad < 6) printf("mismatch it=%d mode=%d len=%u start=%u\n", it, mode, len, start); }
    }
    printf("affine tree: %s\n", bad ? "FAIL" : "ok");
    return bad ? 1 : 0;
}
