// h2mo.cu -- fourth translation unit of libb200zk.so: the flows the MSM / NTT / Fr-vector kernels were built for,
// composed inside the library so that a host shim makes ONE call per opening / per batch of proofs:
//
//   * the Fiat-Shamir transcript of the reference (CardanoFriendlyBlake2b: unkeyed blake2b-256 over the whole absorbed
//     history, /root/reference/src/plutus_gen/adjusted_types/mod.rs:30-72; verifier twin
//     /root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:19-106);
//   * KZGCommitmentScheme::multi_open -- the halo2 multi-open ("H2MO") prover with every polynomial resident in HBM;
//     message order X1, X2, FCommitment, X3, QEvals.., X4, PI as pinned by
//     /root/reference/src/plutus_gen/extraction/pcs/kzg.rs:55-79;
//   * KZGCommitmentScheme::multi_prepare -- the verifier's scalar side and its DualMSM guard
//     (/root/reference/aiken-verifier/aiken_halo2/lib/halo2_kzg.ak:15-171,
//     /root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/Halo2MultiOpenMSM.hs:60-189);
//   * Guard::verify / batch_verify up to the pairing: decompression of every proof point on the GPU and the two final
//     sums, split over the bound GPUs (/root/reference/src/circuits/schnorr_circuit.rs:224-229).
//
// Scalar-side arithmetic runs on the host with the very field code of the kernels (field.cuh is __host__ __device__); it is
// O(queries) per proof.  Everything O(n) runs on the device through the library's own entry points.  The pairing itself
// stays with the caller's pairing library.  No CPU fallback for the device parts.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "b200zk.h"
#include "ctx.hpp"
#include "field.cuh"

using b200zk::Fr;
using b200zk::FrParams;
namespace ctx = b200zk_ctx;

namespace {

// ------------------------------------------------------------------------------------------
// blake2b-256, unkeyed (RFC 7693)
// ------------------------------------------------------------------------------------------
const uint64_t B2_IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                           0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
const uint8_t B2_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
void b2_compress(uint64_t h[8], const uint8_t block[128], uint64_t t, bool last) {
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) memcpy(&m[i], block + 8 * i, 8);   // little-endian host
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2_IV[i]; }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
    auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
        v[a] = v[a] + v[b] + x; v[d] = rotr64(v[d] ^ v[a], 32);
        v[c] = v[c] + v[d];     v[b] = rotr64(v[b] ^ v[c], 24);
        v[a] = v[a] + v[b] + y; v[d] = rotr64(v[d] ^ v[a], 16);
        v[c] = v[c] + v[d];     v[b] = rotr64(v[b] ^ v[c], 63);
    };
    for (int r = 0; r < 12; r++) {
        const uint8_t* s = B2_SIGMA[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
void blake2b_256(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint64_t h[8];
    for (int i = 0; i < 8; i++) h[i] = B2_IV[i];
    h[0] ^= 0x01010000ull ^ 32ull;          // digest length 32, no key, fanout = depth = 1
    uint8_t block[128];
    size_t off = 0;
    while (len - off > 128) {
        b2_compress(h, data + off, (uint64_t)(off + 128), false);
        off += 128;
    }
    memset(block, 0, 128);
    if (len > off) memcpy(block, data + off, len - off);
    b2_compress(h, block, (uint64_t)len, true);
    memcpy(out, h, 32);
}

// ------------------------------------------------------------------------------------------
// host-side Fr (the kernels' Montgomery code, carry flag emulated on the host)
// ------------------------------------------------------------------------------------------
Fr mul(const Fr& a, const Fr& b);
Fr fr_load(const uint8_t b[32]) {                 // canonical little-endian, values >= r reduced like the wire format does
    Fr v, r2;
    memcpy(v.l, b, 32);
    b200zk::fe_reduce_loose(v);
    for (int i = 0; i < 8; i++) r2.l[i] = FrParams::r2(i);
    return mul(v, r2);
}
void fr_store(const Fr& m, uint8_t out[32]) {
    Fr one = b200zk::fe_zero<FrParams>();
    one.l[0] = 1;
    Fr v = mul(m, one);
    memcpy(out, v.l, 32);
}
Fr fr_zero() { return b200zk::fe_zero<FrParams>(); }
Fr fr_one() { return b200zk::fe_one<FrParams>(); }
// Montgomery product on 64-bit limbs with 128-bit intermediates (same value as the kernels' 32-bit-limb fe_mul, whose
// host build emulates the carry flag and is ~20x slower): a batch of 1 024 guards scales ~27 000 terms on the host.
Fr mul(const Fr& a, const Fr& b) {
    typedef unsigned __int128 u128;
    static const uint64_t P[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
    static const uint64_t INV = 0xfffffffeffffffffull;          // -r^-1 mod 2^64
    uint64_t x[4], y[4], t[6] = {0, 0, 0, 0, 0, 0};
    memcpy(x, a.l, 32);
    memcpy(y, b.l, 32);
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)x[j] * y[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * INV;
        c = (u128)m * P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    // t < 2r: one conditional subtraction
    uint64_t d[4];
    u128 bw = 0;
    for (int j = 0; j < 4; j++) {
        u128 v = (u128)t[j] - P[j] - (uint64_t)bw;
        d[j] = (uint64_t)v;
        bw = (v >> 64) & 1;
    }
    bool ge = t[4] != 0 || bw == 0;
    Fr r;
    memcpy(r.l, ge ? d : t, 32);
    return r;
}
Fr add(const Fr& a, const Fr& b) { return b200zk::fe_add(a, b); }
Fr sub(const Fr& a, const Fr& b) { return b200zk::fe_sub(a, b); }
Fr neg(const Fr& a) { return b200zk::fe_neg(a); }
Fr inv(const Fr& a) {   // a^(r-2), square and multiply over the fast host product
    Fr acc = fr_one();
    for (int i = 255; i >= 0; i--) {
        acc = mul(acc, acc);
        if ((FrParams::pm2(i >> 5) >> (i & 31)) & 1) acc = mul(acc, a);
    }
    return acc;
}
bool eq(const Fr& a, const Fr& b) { return b200zk::fe_eq(a, b); }
std::vector<Fr> powers(const Fr& x, size_t n) {
    std::vector<Fr> p(n);
    Fr acc = fr_one();
    for (size_t i = 0; i < n; i++) { p[i] = acc; acc = mul(acc, x); }
    return p;
}

// ------------------------------------------------------------------------------------------
// transcript
// ------------------------------------------------------------------------------------------
struct Transcript {
    std::vector<uint8_t> buf;
    void common_scalar(const uint8_t s[32]) {
        buf.push_back(0x01);
        buf.insert(buf.end(), s, s + 32);
    }
    void common_point(const uint8_t c[48]) {
        buf.push_back(0x01);
        buf.insert(buf.end(), c, c + 48);
    }
    // challenge = LE(H) + LE(H(H)) * 2^256 mod r, H = blake2b-256(history || 0x00); the 0x00 stays in the history
    Fr squeeze() {
        buf.push_back(0x00);
        uint8_t h[32], hh[32];
        blake2b_256(buf.data(), buf.size(), h);
        blake2b_256(h, 32, hh);
        Fr r2;
        for (int i = 0; i < 8; i++) r2.l[i] = FrParams::r2(i);
        return add(fr_load(h), mul(fr_load(hh), r2));   // to_mont(R mod r) = R^2 mod r
    }
};

std::mutex g_mu;
std::map<uint64_t, std::unique_ptr<Transcript>> g_transcripts;
uint64_t g_next = 1;

Transcript* find_transcript(uint64_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_transcripts.find(h);
    return it == g_transcripts.end() ? nullptr : it->second.get();
}

// ------------------------------------------------------------------------------------------
// rotation sets: what precompute_intermediate_sets does with the prover's / verifier's queries
// (/root/reference/src/plutus_gen/extraction/pcs/mod.rs:36-109): commitments in order of first appearance, each with the
// set of points it is opened at; equal point sets share an index, numbered in order of first appearance.  Inside a set the
// points are kept in canonical-value order (the reference orders symbolic rotations; nothing the proof contains depends on
// the order inside a set, only prover and verifier must agree).
// ------------------------------------------------------------------------------------------
struct PointKey {
    uint8_t b[32];
    bool operator<(const PointKey& o) const {
        for (int i = 31; i >= 0; i--)
            if (b[i] != o.b[i]) return b[i] < o.b[i];
        return false;
    }
    bool operator==(const PointKey& o) const { return memcmp(b, o.b, 32) == 0; }
};
struct Sets {
    std::vector<std::vector<PointKey>> points;        // per set, sorted
    std::vector<std::vector<uint32_t>> members;       // per set: polynomial / commitment indices, in order of first appearance
    std::vector<uint32_t> set_of;                     // per polynomial
    std::vector<std::vector<Fr>> evals;               // per polynomial: value at each point of its set (verifier side)
};
int32_t build_sets(uint32_t n_polys, const uint32_t* q_poly, const uint8_t* q_points, const uint8_t* q_evals, uint32_t n_queries, Sets& out) {
    std::vector<std::vector<std::pair<PointKey, Fr>>> per(n_polys);
    std::vector<uint32_t> order;
    std::vector<uint8_t> seen(n_polys, 0);
    for (uint32_t i = 0; i < n_queries; i++) {
        uint32_t p = q_poly[i];
        if (p >= n_polys) return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-open: a query names a polynomial that does not exist");
        PointKey k;
        Fr canon;
        memcpy(canon.l, q_points + 32 * (size_t)i, 32);
        b200zk::fe_reduce_loose(canon);
        memcpy(k.b, canon.l, 32);
        Fr ev = q_evals ? fr_load(q_evals + 32 * (size_t)i) : fr_zero();
        bool dup = false;
        for (auto& e : per[p])
            if (e.first == k) {
                if (q_evals && !eq(e.second, ev)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-open: two evaluations for the same polynomial and point");
                dup = true;
            }
        if (!dup) per[p].push_back({k, ev});
        if (!seen[p]) { seen[p] = 1; order.push_back(p); }
    }
    out.set_of.assign(n_polys, 0xffffffffu);
    out.evals.assign(n_polys, {});
    for (uint32_t p : order) {
        auto& v = per[p];
        std::sort(v.begin(), v.end(), [](const std::pair<PointKey, Fr>& a, const std::pair<PointKey, Fr>& b) { return a.first < b.first; });
        std::vector<PointKey> keys;
        for (auto& e : v) keys.push_back(e.first);
        uint32_t s = 0;
        for (; s < out.points.size(); s++)
            if (out.points[s] == keys) break;
        if (s == out.points.size()) { out.points.push_back(keys); out.members.emplace_back(); }
        out.members[s].push_back(p);
        out.set_of[p] = s;
        for (auto& e : v) out.evals[p].push_back(e.second);
    }
    return B200ZK_OK;
}

// coefficients (low to high) of the polynomial through (x_i, y_i), by Newton's divided differences; |set| is tiny
std::vector<Fr> interpolate(const std::vector<Fr>& xs, const std::vector<Fr>& ys) {
    size_t m = xs.size();
    std::vector<Fr> dd = ys;
    for (size_t j = 1; j < m; j++)
        for (size_t i = m - 1; i >= j; i--) dd[i] = mul(sub(dd[i], dd[i - 1]), inv(sub(xs[i], xs[i - j])));
    std::vector<Fr> coef(m, fr_zero());   // Horner over the Newton basis
    for (size_t k = m; k-- > 0;) {
        // coef = coef * (X - xs[k]) + dd[k]
        for (size_t d = m - 1; d > 0; d--) coef[d] = sub(coef[d - 1], mul(coef[d], xs[k]));
        coef[0] = add(neg(mul(coef[0], xs[k])), dd[k]);
    }
    return coef;
}
Fr eval_poly(const std::vector<Fr>& c, const Fr& x) {
    Fr acc = fr_zero();
    for (size_t i = c.size(); i-- > 0;) acc = add(mul(acc, x), c[i]);
    return acc;
}

// the verifier's scalar pipeline (halo2_kzg.ak:46-171): q_eval_sets, f_eval, v
struct H2moScalars {
    std::vector<std::vector<Fr>> q_eval_sets;   // per set, per point
    Fr f_eval, v;
};
H2moScalars h2mo_scalars(const Sets& sets, const Fr& x1, const Fr& x2, const Fr& x3, const Fr& x4, const std::vector<Fr>& proof_q_evals) {
    H2moScalars o;
    const size_t S = sets.points.size();
    o.q_eval_sets.resize(S);
    for (size_t s = 0; s < S; s++) {
        std::vector<Fr> acc(sets.points[s].size(), fr_zero());
        Fr xp = fr_one();
        for (uint32_t p : sets.members[s]) {
            for (size_t k = 0; k < acc.size(); k++) acc[k] = add(acc[k], mul(sets.evals[p][k], xp));
            xp = mul(xp, x1);
        }
        o.q_eval_sets[s] = acc;
    }
    // f_eval = sum_s x2^s (q_s(x3) - r_s(x3)) / prod_{p in set s} (x3 - p)   (the fold of compute_f_eval runs over the reversed list)
    Fr acc = fr_zero();
    for (size_t s = S; s-- > 0;) {
        std::vector<Fr> xs;
        for (auto& k : sets.points[s]) { Fr c; memcpy(c.l, k.b, 32); xs.push_back(b200zk::fe_to_mont(c)); }
        Fr r_eval = eval_poly(interpolate(xs, o.q_eval_sets[s]), x3);
        Fr den = fr_one();
        for (auto& x : xs) den = mul(den, sub(x3, x));
        acc = add(mul(acc, x2), mul(sub(proof_q_evals[s], r_eval), inv(den)));
    }
    o.f_eval = acc;
    Fr v = fr_zero(), xp = fr_one();
    for (size_t s = 0; s < S; s++) { v = add(v, mul(xp, proof_q_evals[s])); xp = mul(xp, x4); }
    o.v = add(v, mul(xp, acc));
    return o;
}

// ------------------------------------------------------------------------------------------
// guards: the verifier's pair of lazy MSMs (upstream DualMSM), as (scalar, compressed point) terms
// ------------------------------------------------------------------------------------------
struct Term {
    Fr s;
    uint8_t pt[48];
};
struct Guard {
    std::vector<Term> left, right;
    Fr g_scalar = b200zk::fe_zero<FrParams>();    // coefficient of the generator on the right (kept apart: batches merge it)
};
std::map<uint64_t, std::unique_ptr<Guard>> g_guards;

Guard* find_guard(uint64_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_guards.find(h);
    return it == g_guards.end() ? nullptr : it->second.get();
}

// compressed generator of G1 (/root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:125)
const uint8_t G1_GEN_COMPRESSED[48] = {0x97, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f,
                                       0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58,
                                       0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb};

// sum_i s_i * decompress(pt_i): every point decompressed on the GPU in one batch, then one ad-hoc MSM (split over the bound GPUs)
int32_t eval_terms(const std::vector<Term>& terms, uint8_t out[96]) {
    const size_t n = terms.size();
    if (n == 0) { memset(out, 0, 96); return B200ZK_OK; }
    std::vector<uint8_t> comp(48 * n), aff(96 * n), sc(32 * n);
    for (size_t i = 0; i < n; i++) {
        memcpy(&comp[48 * i], terms[i].pt, 48);
        memcpy(&sc[32 * i], terms[i].s.l, 32);       // Montgomery limbs as they are: the MSM reads them in that form
    }
    XTRY(b200zk_g1_decompress_batch(comp.data(), n, aff.data(), nullptr));
    return b200zk_msm_g1_adhoc(aff.data(), B200ZK_FMT_CANONICAL, sc.data(), B200ZK_FMT_MONT, n, out);
}

// Device scratch of one multi_open: a grow-only arena per concurrent call and device, kept between calls.  cudaMalloc / cudaFree
// inside the flow were measured at tens to hundreds of milliseconds per opening (and one ~0.9 s stall in five calls at k = 19)
// against 8-40 ms for everything else, so the flow allocates nothing once it has run at its largest size.
struct Arena {
    int ordinal = -1;
    uint8_t* p = nullptr;
    size_t cap = 0;
    bool busy = false;
};
std::mutex g_arena_mu;
std::vector<Arena*>* g_arenas = nullptr;   // leaked on purpose (no static destruction order games); emptied by b200zk_shutdown
void arenas_shutdown() {
    std::lock_guard<std::mutex> lk(g_arena_mu);
    if (!g_arenas) return;
    for (Arena* a : *g_arenas) {
        if (a->p) {
            ctx::DeviceScope scope(a->ordinal);
            cudaFree(a->p);
        }
        delete a;
    }
    g_arenas->clear();
}
struct DevMem {
    Arena* a = nullptr;
    size_t used = 0;
    // the caller has made `ordinal` the current CUDA device
    int32_t reserve(int ordinal, size_t bytes) {
        {
            std::lock_guard<std::mutex> lk(g_arena_mu);
            if (!g_arenas) {
                g_arenas = new std::vector<Arena*>();
                ctx::on_shutdown(arenas_shutdown);
            }
            for (Arena* c : *g_arenas)   // a free arena of this device, the largest first choice
                if (!c->busy && c->ordinal == ordinal && (!a || c->cap > a->cap)) a = c;
            if (!a) {
                a = new Arena();
                a->ordinal = ordinal;
                g_arenas->push_back(a);
            }
            a->busy = true;
        }
        if (a->cap < bytes) {
            if (a->p) XCU(cudaFree(a->p));
            a->p = nullptr;
            a->cap = 0;
            XCU(cudaMalloc(&a->p, bytes));
            a->cap = bytes;
        }
        return B200ZK_OK;
    }
    void* take(size_t bytes) {
        void* r = a->p + used;
        used += (bytes + 255) & ~(size_t)255;
        return r;
    }
    ~DevMem() {
        if (!a) return;
        std::lock_guard<std::mutex> lk(g_arena_mu);
        a->busy = false;
    }
};

}  // namespace

extern "C" {

// ---- transcript ------------------------------------------------------------------------------
int32_t b200zk_transcript_new(uint64_t* out_transcript) {
    if (!out_transcript) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(g_mu);
    *out_transcript = g_next++;
    g_transcripts[*out_transcript] = std::make_unique<Transcript>();
    return B200ZK_OK;
}
int32_t b200zk_transcript_free(uint64_t transcript) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_transcripts.erase(transcript)) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown transcript handle");
    return B200ZK_OK;
}
int32_t b200zk_transcript_common_scalar(uint64_t transcript, const uint8_t scalar[32]) {
    Transcript* t = find_transcript(transcript);
    if (!t) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown transcript handle");
    if (!scalar) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    t->common_scalar(scalar);
    return B200ZK_OK;
}
int32_t b200zk_transcript_common_point(uint64_t transcript, const uint8_t compressed[48]) {
    Transcript* t = find_transcript(transcript);
    if (!t) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown transcript handle");
    if (!compressed) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    t->common_point(compressed);
    return B200ZK_OK;
}
int32_t b200zk_transcript_squeeze(uint64_t transcript, uint8_t out_challenge[32]) {
    Transcript* t = find_transcript(transcript);
    if (!t) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown transcript handle");
    if (!out_challenge) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    fr_store(t->squeeze(), out_challenge);
    return B200ZK_OK;
}

// ---- prover: multi_open with resident polynomials ----------------------------------------------
int32_t b200zk_h2mo_open_dev(uint64_t bases, uint64_t transcript, const void* const* d_polys, uint32_t n_polys, uint64_t n,
                             const uint32_t* query_poly, const uint8_t* query_points, uint32_t n_queries, uint8_t* out_proof,
                             size_t cap, size_t* out_len) {
    Transcript* tr = find_transcript(transcript);
    if (!tr) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown transcript handle");
    if (!d_polys || !query_poly || !query_points || !out_proof || !out_len || n_polys == 0 || n_queries == 0 || n == 0)
        return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-open: null pointer or empty query list");
    // B200ZK_H2MO_TIMING=1: phase times on stderr (the device is synchronised at every mark; diagnostics only)
    static const bool timing = getenv("B200ZK_H2MO_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[h2mo_open n=%llu] %-14s %8.3f ms\n", (unsigned long long)n, what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    Sets sets;
    XTRY(build_sets(n_polys, query_poly, query_points, nullptr, n_queries, sets));
    const size_t S = sets.points.size();
    const size_t need = 48 + 32 * S + 48;
    *out_len = need;
    if (cap < need) return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-open: proof buffer too small");
    for (auto& pts : sets.points)
        if (pts.size() >= n) return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-open: more opening points than coefficients");
    ctx::Dev* dev = nullptr;
    XTRY(ctx::for_pointer(d_polys[0], &dev));
    ctx::DeviceScope scope(ctx::ordinal(dev));
    DevMem mem;
    const size_t pb = ((size_t)n * 32 + 255) & ~(size_t)255;
    size_t M = S;                                      // slots of the small staging area: evaluations | interpolation coefficients
    for (auto& pts : sets.points) M = std::max(M, pts.size());
    XTRY(mem.reserve(ctx::ordinal(dev), (S + 3) * pb + 64 * M + 96 + 512));
    std::vector<void*> Q(S);
    for (size_t s = 0; s < S; s++) Q[s] = mem.take(pb);
    void *G = mem.take(pb), *T = mem.take(pb), *F = mem.take(pb), *small = mem.take(64 * M), *d_pt = mem.take(96);
    uint8_t tmp[32];
    auto canon = [&](const Fr& x, uint8_t out[32]) { fr_store(x, out); };
    mark("sets+alloc");

    // x1: q_s = sum_j x1^j p_{s,j}
    Fr x1 = tr->squeeze();
    for (size_t s = 0; s < S; s++) {
        std::vector<const void*> ps;
        std::vector<uint8_t> cf;
        Fr xp = fr_one();
        for (uint32_t p : sets.members[s]) {
            if (!d_polys[p]) return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-open: null polynomial");
            ps.push_back(d_polys[p]);
            canon(xp, tmp);
            cf.insert(cf.end(), tmp, tmp + 32);
            xp = mul(xp, x1);
        }
        XTRY(b200zk_fr_lincomb_dev(ps.data(), cf.data(), (uint32_t)ps.size(), Q[s], n, nullptr));
    }
    mark("q_s lincombs");
    // x2: f = sum_s x2^s (q_s - r_s) / Z_s, r_s the interpolation of q_s on its point set, Z_s = prod (X - point)
    Fr x2 = tr->squeeze();
    std::vector<Fr> x2p = powers(x2, S);
    XCU(cudaMemsetAsync(F, 0, pb, nullptr));
    for (size_t s = 0; s < S; s++) {
        const size_t m = sets.points[s].size();
        std::vector<Fr> xs, ys(m);
        std::vector<uint8_t> evb(32 * m);
        for (size_t k = 0; k < m; k++) {
            XTRY(b200zk_fr_kate_div_dev(Q[s], n, sets.points[s][k].b, nullptr, (uint8_t*)small + 32 * k, nullptr));
            Fr c;
            memcpy(c.l, sets.points[s][k].b, 32);
            xs.push_back(b200zk::fe_to_mont(c));
        }
        XTRY(b200zk_dev_download(evb.data(), small, 32 * m));     // Montgomery limbs
        for (size_t k = 0; k < m; k++) memcpy(ys[k].l, &evb[32 * k], 32);
        std::vector<Fr> r = interpolate(xs, ys);
        // G = q_s with r_s subtracted from its low coefficients, then divided by every (X - point): exact divisions
        XCU(cudaMemcpyAsync(G, Q[s], pb, cudaMemcpyDeviceToDevice, nullptr));
        std::vector<uint8_t> rb(32 * m);
        for (size_t k = 0; k < m; k++) memcpy(&rb[32 * k], r[k].l, 32);
        XTRY(b200zk_dev_upload((uint8_t*)small + 32 * M, rb.data(), 32 * m));
        XTRY(b200zk_fr_pointwise_dev(2, G, (uint8_t*)small + 32 * M, nullptr, G, m, nullptr));
        uint64_t len = n;
        void *src = G, *dst = T;
        for (size_t k = 0; k < m; k++) {
            XTRY(b200zk_fr_kate_div_dev(src, len, sets.points[s][k].b, dst, nullptr, nullptr));
            len--;
            std::swap(src, dst);
        }
        XCU(cudaMemsetAsync((uint8_t*)src + 32 * len, 0, 32 * (n - len), nullptr));
        canon(x2p[s], tmp);
        // F += x2^s * src   (scale into the other buffer, then add)
        XTRY(b200zk_fr_pointwise_dev(3, src, nullptr, tmp, dst, n, nullptr));
        XTRY(b200zk_fr_pointwise_dev(1, F, dst, nullptr, F, n, nullptr));
    }
    mark("f polynomial");
    // f commitment
    uint8_t aff[96], comp[48];
    XTRY(b200zk_msm_g1_dev(bases, 0, F, n, 1, B200ZK_FMT_MONT, nullptr, d_pt, nullptr));
    XTRY(b200zk_dev_download(aff, d_pt, 96));
    mark("commit f");
    XTRY(b200zk_g1_compress(aff, comp));
    tr->common_point(comp);
    memcpy(out_proof, comp, 48);
    // x3: q_s(x3)
    Fr x3 = tr->squeeze();
    uint8_t x3b[32];
    canon(x3, x3b);
    for (size_t s = 0; s < S; s++) XTRY(b200zk_fr_kate_div_dev(Q[s], n, x3b, nullptr, (uint8_t*)small + 32 * s, nullptr));
    {
        std::vector<uint8_t> evb(32 * S);
        XTRY(b200zk_dev_download(evb.data(), small, 32 * S));
        for (size_t s = 0; s < S; s++) {
            Fr e;
            memcpy(e.l, &evb[32 * s], 32);
            fr_store(e, out_proof + 48 + 32 * s);
            tr->common_scalar(out_proof + 48 + 32 * s);
        }
    }
    mark("q_s(x3)");
    // x4: final = sum_s x4^s q_s + x4^S f; pi commits to (final - final(x3)) / (X - x3)
    Fr x4 = tr->squeeze();
    {
        std::vector<const void*> ps(Q.begin(), Q.end());
        ps.push_back(F);
        std::vector<Fr> x4p = powers(x4, S + 1);
        std::vector<uint8_t> cf(32 * (S + 1));
        for (size_t s = 0; s <= S; s++) canon(x4p[s], &cf[32 * s]);
        XTRY(b200zk_fr_lincomb_dev(ps.data(), cf.data(), (uint32_t)ps.size(), G, n, nullptr));
    }
    XTRY(b200zk_fr_kate_div_dev(G, n, x3b, T, nullptr, nullptr));
    mark("final / (X-x3)");
    XTRY(b200zk_msm_g1_dev(bases, 0, T, n - 1, 1, B200ZK_FMT_MONT, nullptr, d_pt, nullptr));
    XTRY(b200zk_dev_download(aff, d_pt, 96));
    mark("commit pi");
    XTRY(b200zk_g1_compress(aff, comp));
    tr->common_point(comp);
    memcpy(out_proof + 48 + 32 * S, comp, 48);
    return B200ZK_OK;
}

// ---- verifier: multi_prepare --------------------------------------------------------------------
int32_t b200zk_h2mo_prepare(uint64_t transcript, const uint8_t* commitments, uint32_t n_commitments, const uint32_t* query_commitment,
                            const uint8_t* query_points, const uint8_t* query_evals, uint32_t n_queries, const uint8_t* proof,
                            size_t proof_len, uint64_t* out_guard, uint8_t* out_scalars) {
    Transcript* tr = find_transcript(transcript);
    if (!tr) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown transcript handle");
    if (!commitments || !query_commitment || !query_points || !query_evals || !proof || !out_guard || n_commitments == 0 || n_queries == 0)
        return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-prepare: null pointer or empty query list");
    Sets sets;
    XTRY(build_sets(n_commitments, query_commitment, query_points, query_evals, n_queries, sets));
    const size_t S = sets.points.size();
    if (proof_len != 48 + 32 * S + 48) return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-prepare: proof has the wrong length for these queries");
    Fr x1 = tr->squeeze();
    Fr x2 = tr->squeeze();
    const uint8_t* f_comm = proof;
    tr->common_point(f_comm);
    Fr x3 = tr->squeeze();
    std::vector<Fr> q_evals(S);
    for (size_t s = 0; s < S; s++) {
        // a scalar on the wire must be canonical (< r): anything else is a malformed proof
        Fr raw, red;
        memcpy(raw.l, proof + 48 + 32 * s, 32);
        red = raw;
        b200zk::fe_reduce_loose(red);
        if (!eq(raw, red)) return ctx::fail(B200ZK_ERR_INVALID_ARG, "multi-prepare: non-canonical scalar in the proof");
        q_evals[s] = fr_load(proof + 48 + 32 * s);
        tr->common_scalar(proof + 48 + 32 * s);
    }
    Fr x4 = tr->squeeze();
    const uint8_t* pi = proof + 48 + 32 * S;
    tr->common_point(pi);
    H2moScalars sc = h2mo_scalars(sets, x1, x2, x3, x4, q_evals);
    // right = sum_s x4^s sum_j x1^j C_{s,j} + x4^S f_commitment - v G + x3 pi ; left = pi   (halo2_kzg.ak:15-44)
    auto g = std::make_unique<Guard>();
    std::vector<Fr> x4p = powers(x4, S + 1);
    for (size_t s = 0; s < S; s++) {
        Fr xp = x4p[s];
        for (uint32_t c : sets.members[s]) {
            Term t;
            t.s = xp;
            memcpy(t.pt, commitments + 48 * (size_t)c, 48);
            g->right.push_back(t);
            xp = mul(xp, x1);
        }
    }
    Term tf, tp, tl;
    tf.s = x4p[S];
    memcpy(tf.pt, f_comm, 48);
    g->right.push_back(tf);
    tp.s = x3;
    memcpy(tp.pt, pi, 48);
    g->right.push_back(tp);
    g->g_scalar = neg(sc.v);
    tl.s = fr_one();
    memcpy(tl.pt, pi, 48);
    g->left.push_back(tl);
    if (out_scalars) {   // x1, x2, x3, x4, f_eval, v: what the reference's own tests pin (Halo2MultiOpenMSM.hs:26-42)
        fr_store(x1, out_scalars);
        fr_store(x2, out_scalars + 32);
        fr_store(x3, out_scalars + 64);
        fr_store(x4, out_scalars + 96);
        fr_store(sc.f_eval, out_scalars + 128);
        fr_store(sc.v, out_scalars + 160);
    }
    std::lock_guard<std::mutex> lk(g_mu);
    *out_guard = g_next++;
    g_guards[*out_guard] = std::move(g);
    return B200ZK_OK;
}

// the scalar pipeline alone, with the challenges given: the known-answer surface of the reference's tests
int32_t b200zk_h2mo_scalars(uint32_t n_commitments, const uint32_t* query_commitment, const uint8_t* query_points, const uint8_t* query_evals,
                            uint32_t n_queries, const uint8_t challenges[128], const uint8_t* proof_q_evals, uint32_t n_sets,
                            uint8_t* out_q_eval_sets, size_t cap_q_eval_sets, uint8_t out_f_eval[32], uint8_t out_v[32]) {
    if (!query_commitment || !query_points || !query_evals || !challenges || !proof_q_evals || !out_f_eval || !out_v)
        return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer");
    Sets sets;
    XTRY(build_sets(n_commitments, query_commitment, query_points, query_evals, n_queries, sets));
    if (sets.points.size() != n_sets) return ctx::fail(B200ZK_ERR_INVALID_ARG, "h2mo: the queries form a different number of point sets");
    std::vector<Fr> pq(n_sets);
    for (uint32_t s = 0; s < n_sets; s++) pq[s] = fr_load(proof_q_evals + 32 * (size_t)s);
    H2moScalars sc = h2mo_scalars(sets, fr_load(challenges), fr_load(challenges + 32), fr_load(challenges + 64), fr_load(challenges + 96), pq);
    fr_store(sc.f_eval, out_f_eval);
    fr_store(sc.v, out_v);
    if (out_q_eval_sets) {
        size_t off = 0;
        for (auto& set : sc.q_eval_sets)
            for (auto& e : set) {
                if (off + 32 > cap_q_eval_sets) return ctx::fail(B200ZK_ERR_INVALID_ARG, "h2mo: q_eval_sets buffer too small");
                fr_store(e, out_q_eval_sets + off);
                off += 32;
            }
    }
    return B200ZK_OK;
}

int32_t b200zk_guard_free(uint64_t guard) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_guards.erase(guard)) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown guard handle");
    return B200ZK_OK;
}

// left and right sums of sum_i c_i * guard_i (c = NULL: a single guard, c_0 = 1): Guard::verify / batch_verify without the pairing
int32_t b200zk_guard_eval(const uint64_t* guards, uint32_t n_guards, const uint8_t* challenges, uint8_t out_left[96], uint8_t out_right[96]) {
    if (!guards || !out_left || !out_right || n_guards == 0) return ctx::fail(B200ZK_ERR_INVALID_ARG, "null pointer or no guard");
    std::vector<Term> left, right;
    {
        size_t nl = 0, nr = 1;
        for (uint32_t i = 0; i < n_guards; i++) {
            Guard* g = find_guard(guards[i]);
            if (!g) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown guard handle");
            nl += g->left.size();
            nr += g->right.size();
        }
        left.reserve(nl);
        right.reserve(nr);
    }
    Fr gsum = fr_zero();
    for (uint32_t i = 0; i < n_guards; i++) {
        Guard* g = find_guard(guards[i]);
        if (!g) return ctx::fail(B200ZK_ERR_BAD_HANDLE, "unknown guard handle");
        Fr c = challenges ? fr_load(challenges + 32 * (size_t)i) : fr_one();
        for (const Term& t : g->left) { Term u = t; u.s = mul(t.s, c); left.push_back(u); }
        for (const Term& t : g->right) { Term u = t; u.s = mul(t.s, c); right.push_back(u); }
        gsum = add(gsum, mul(g->g_scalar, c));
    }
    Term tg;
    tg.s = gsum;
    memcpy(tg.pt, G1_GEN_COMPRESSED, 48);
    right.push_back(tg);
    XTRY(eval_terms(left, out_left));
    return eval_terms(right, out_right);
}

}  // extern "C"
