// b200zk.hpp -- header-only C++ mirror of the reference-facing interface for the MSM / NTT path,
// layered on the C ABI of b200zk.h.  It marshals buffers and nothing else (no arithmetic, no
// fallback).  Names and argument meaning follow the upstream traits the reference reaches:
//   ParamsKZG                        /root/reference/src/kzg_params.rs:33-80
//   KZGCommitmentScheme::commit*     /root/reference/examples/simple_mul.rs:62,72 (via keygen_vk / create_proof)
//   EvaluationDomain                 /root/reference/examples/ivc.rs:109, /root/reference/src/circuits/ivc_circuit.rs:305
//   DualMSM (Guard::verify)          /root/reference/examples/simple_mul.rs:98-102
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b200zk.h"

namespace b200zk {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int32_t rc) {
    if (rc == B200ZK_OK) return;
    char buf[1024] = {0};
    b200zk_last_error(buf, sizeof buf);
    throw Error(rc, buf);
}

using Fr = std::array<uint8_t, 32>;        // little-endian canonical
using G1Affine = std::array<uint8_t, 96>;  // x || y little-endian canonical, identity = (0,0)

inline void init(int device = -1) { check(b200zk_init(device)); }
// One process, several GPUs (the reference prover is a single process, /root/reference/examples/simple_mul.rs:39-141): every
// later commit / transform fans out over all of them.  Empty list: the first n_devices ordinals.
inline void init_devices(const std::vector<int32_t>& ordinals, int32_t n_devices = 0) {
    check(ordinals.empty() ? b200zk_init_devices(nullptr, n_devices) : b200zk_init_devices(ordinals.data(), (int32_t)ordinals.size()));
}

// Device residency of the two SRS tables (g: monomial basis, g_lagrange: Lagrange basis).
class ParamsKZG {
public:
    ParamsKZG(uint32_t k, const uint8_t* g, const uint8_t* g_lagrange, uint32_t fmt = B200ZK_FMT_CANONICAL,
              uint32_t stride = 96)
        : k_(k), n_(uint64_t(1) << k) {
        check(b200zk_bases_register(g, n_, fmt, stride, &g_));
        if (g_lagrange) check(b200zk_bases_register(g_lagrange, n_, fmt, stride, &gl_));
    }
    // ParamsKZG::unsafe_setup (/root/reference/src/kzg_params.rs:43): both tables generated on the device for the secret s
    // (canonical Fr) and the domain generator omega of order 2^k, and registered without leaving HBM.
    static ParamsKZG unsafe_setup(uint32_t k, const std::array<uint8_t, 32>& s, const std::array<uint8_t, 32>& omega) {
        ParamsKZG p(k);
        void *g = nullptr, *gl = nullptr;
        check(b200zk_dev_alloc(&g, p.n_ * 96));
        check(b200zk_dev_alloc(&gl, p.n_ * 96));
        int32_t rc = b200zk_srs_generate_dev(s.data(), k, omega.data(), g, gl, nullptr);
        if (rc == B200ZK_OK) rc = b200zk_bases_register_dev(g, p.n_, B200ZK_FMT_MONT, 96, &p.g_);
        if (rc == B200ZK_OK) rc = b200zk_bases_register_dev(gl, p.n_, B200ZK_FMT_MONT, 96, &p.gl_);
        b200zk_dev_free(g);
        b200zk_dev_free(gl);
        check(rc);
        return p;
    }
    ParamsKZG(ParamsKZG&& o) noexcept : k_(o.k_), n_(o.n_), g_(o.g_), gl_(o.gl_) { o.g_ = o.gl_ = 0; }
    ~ParamsKZG() {
        if (g_) b200zk_bases_release(g_);
        if (gl_) b200zk_bases_release(gl_);
    }
    ParamsKZG(const ParamsKZG&) = delete;
    ParamsKZG& operator=(const ParamsKZG&) = delete;
    uint32_t k() const { return k_; }
    uint64_t n() const { return n_; }
    uint64_t g() const { return g_; }
    uint64_t g_lagrange() const { return gl_; }

private:
    explicit ParamsKZG(uint32_t k) : k_(k), n_(uint64_t(1) << k) {}
    uint32_t k_;
    uint64_t n_, g_ = 0, gl_ = 0;
};

struct KZGCommitmentScheme {
    // commit to a polynomial in coefficient form: MSM against params.g
    static G1Affine commit(const ParamsKZG& params, const std::vector<Fr>& poly) { return msm(params.g(), params.n(), poly); }
    // commit to a polynomial in Lagrange form: MSM against params.g_lagrange
    static G1Affine commit_lagrange(const ParamsKZG& params, const std::vector<Fr>& poly) {
        if (!params.g_lagrange()) throw Error(B200ZK_ERR_INVALID_ARG, "params built without a Lagrange-basis table");
        return msm(params.g_lagrange(), params.n(), poly);
    }
    // all columns of one prover phase in a single call; every polynomial stays in its own vector (pointer-array entry point),
    // with several GPUs bound the columns are dealt out or, for a partitioned table, every column is sharded by point range
    static std::vector<G1Affine> commit_batch(const ParamsKZG& params, const std::vector<const std::vector<Fr>*>& polys, bool lagrange = false) {
        std::vector<G1Affine> out(polys.size());
        if (polys.empty()) return out;
        const uint64_t h = lagrange ? params.g_lagrange() : params.g();
        if (!h) throw Error(B200ZK_ERR_INVALID_ARG, "params built without a Lagrange-basis table");
        std::vector<const uint8_t*> ptrs;
        for (auto* p : polys) {
            if (p->size() != polys[0]->size() || p->size() > params.n()) throw Error(B200ZK_ERR_INVALID_ARG, "columns of one batch have the same length, at most the SRS size");
            ptrs.push_back(reinterpret_cast<const uint8_t*>(p->data()));
        }
        check(b200zk_msm_g1_batch_ptrs(h, 0, ptrs.data(), polys[0]->size(), (uint32_t)polys.size(), B200ZK_FMT_CANONICAL,
                                       reinterpret_cast<uint8_t*>(out.data())));
        return out;
    }

private:
    static G1Affine msm(uint64_t handle, uint64_t n_max, const std::vector<Fr>& poly) {
        if (poly.size() > n_max) throw Error(B200ZK_ERR_INVALID_ARG, "polynomial longer than the SRS");
        G1Affine out{};
        check(b200zk_msm_g1(handle, 0, reinterpret_cast<const uint8_t*>(poly.data()), poly.size(), B200ZK_FMT_CANONICAL,
                            out.data()));
        return out;
    }
};

// The four transforms of the halo2 evaluation domain.  Domain constants (omega, its inverse, the
// extended omega and the coset generator) are Fr values the caller already has (upstream computes
// them in EvaluationDomain::new); they are passed in canonical form.
class EvaluationDomain {
public:
    EvaluationDomain(uint32_t k, uint32_t extended_k, Fr omega, Fr omega_inv, Fr extended_omega, Fr extended_omega_inv,
                     Fr g_coset, Fr g_coset_inv, uint32_t quotient_poly_degree)
        : k_(k), ek_(extended_k), w_(omega), wi_(omega_inv), ew_(extended_omega), ewi_(extended_omega_inv), g_(g_coset),
          gi_(g_coset_inv), qd_(quotient_poly_degree) {}

    void lagrange_to_coeff(std::vector<Fr>& a) const { run(a, k_, wi_, B200ZK_NTT_INVERSE_SCALE, nullptr); }
    void coeff_to_lagrange(std::vector<Fr>& a) const { run(a, k_, w_, 0, nullptr); }
    void coeff_to_extended(std::vector<Fr>& a) const {
        a.resize(size_t(1) << ek_, Fr{});
        run(a, ek_, ew_, B200ZK_NTT_COSET_IN, g_.data());
    }
    void extended_to_coeff(std::vector<Fr>& a) const {
        run(a, ek_, ewi_, B200ZK_NTT_INVERSE_SCALE | B200ZK_NTT_COSET_OUT, gi_.data());
        a.resize((size_t(1) << k_) * qd_);
    }

private:
    static void run(std::vector<Fr>& a, uint32_t log_n, const Fr& omega, uint32_t flags, const uint8_t* shift) {
        if (a.size() != (size_t(1) << log_n)) throw Error(B200ZK_ERR_INVALID_ARG, "polynomial length is not the domain size");
        check(b200zk_ntt_fr(reinterpret_cast<uint8_t*>(a.data()), log_n, omega.data(), flags, shift));
    }
    uint32_t k_, ek_;
    Fr w_, wi_, ew_, ewi_, g_, gi_;
    uint32_t qd_;
};

// The verifier's pair of lazy MSMs; eval() returns (left, right).  The accept decision
// e(left, [s]G2) == e(right, G2) (/root/reference/aiken-verifier/templates/verification_h2.hbs:121-128)
// stays with the caller's pairing library.
class DualMSM {
public:
    void append_left(const Fr& s, const G1Affine& p) { ls_.push_back(s); lp_.push_back(p); }
    void append_right(const Fr& s, const G1Affine& p) { rs_.push_back(s); rp_.push_back(p); }
    std::pair<G1Affine, G1Affine> eval() const { return {one(ls_, lp_), one(rs_, rp_)}; }

private:
    static G1Affine one(const std::vector<Fr>& s, const std::vector<G1Affine>& p) {
        G1Affine out{};
        if (s.empty()) return out;
        check(b200zk_msm_g1_adhoc(reinterpret_cast<const uint8_t*>(p.data()), B200ZK_FMT_CANONICAL,
                                  reinterpret_cast<const uint8_t*>(s.data()), B200ZK_FMT_CANONICAL, s.size(), out.data()));
        return out;
    }
    std::vector<Fr> ls_, rs_;
    std::vector<G1Affine> lp_, rp_;
};


// ---------------------------------------------------------------------------------------------------
// Columns resident in HBM (INTEGRATION.md route C): RAII device vectors of Fr in Montgomery
// form and the polynomial-side operations between the NTTs and the commitments.
// ---------------------------------------------------------------------------------------------------
class DeviceFr {
public:
    explicit DeviceFr(uint64_t n) : n_(n) { check(b200zk_dev_alloc(&p_, (n ? n : 1) * 32)); }
    // canonical host values -> Montgomery form on the device
    explicit DeviceFr(const std::vector<Fr>& v) : DeviceFr(v.size()) {
        if (n_) {
            check(b200zk_dev_upload(p_, v.data(), n_ * 32));
            check(b200zk_fr_convert_dev(p_, p_, n_, 1, nullptr));
        }
    }
    ~DeviceFr() { if (p_) b200zk_dev_free(p_); }
    DeviceFr(const DeviceFr&) = delete;
    DeviceFr& operator=(const DeviceFr&) = delete;
    DeviceFr(DeviceFr&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    void* ptr() const { return p_; }
    uint64_t size() const { return n_; }
    std::vector<Fr> to_host() const {   // canonical
        std::vector<Fr> out(n_);
        if (!n_) return out;
        DeviceFr tmp(n_);
        check(b200zk_fr_convert_dev(p_, tmp.p_, n_, 0, nullptr));
        check(b200zk_dev_download(out.data(), tmp.p_, n_ * 32));
        return out;
    }

private:
    void* p_ = nullptr;
    uint64_t n_ = 0;
};

namespace poly {
// (q, p(z)) with p(X) - p(z) = q(X) (X - z): the witness polynomials of the KZG multi-open
// (/root/reference/src/plutus_gen/extraction/pcs/kzg.rs:55-79)
inline std::pair<DeviceFr, Fr> kate_div(const DeviceFr& p, const Fr& z) {
    DeviceFr q(p.size() ? p.size() - 1 : 0), ev(1);
    check(b200zk_fr_kate_div_dev(p.ptr(), p.size(), z.data(), q.ptr(), ev.ptr(), nullptr));
    return {std::move(q), ev.to_host()[0]};
}
inline Fr eval(const DeviceFr& p, const Fr& z) {
    DeviceFr ev(1);
    check(b200zk_fr_kate_div_dev(p.ptr(), p.size(), z.data(), nullptr, ev.ptr(), nullptr));
    return ev.to_host()[0];
}
// sum_k coeffs[k] * polys[k]
inline DeviceFr lincomb(const std::vector<const DeviceFr*>& polys, const std::vector<Fr>& coeffs) {
    if (polys.empty() || polys.size() != coeffs.size()) throw Error(B200ZK_ERR_INVALID_ARG, "lincomb: polys and coeffs must match");
    DeviceFr out(polys[0]->size());
    std::vector<const void*> ptrs;
    for (auto* q : polys) ptrs.push_back(q->ptr());
    check(b200zk_fr_lincomb_dev(ptrs.data(), reinterpret_cast<const uint8_t*>(coeffs.data()), (uint32_t)polys.size(), out.ptr(),
                                out.size(), nullptr));
    return out;
}
// z_0 = 1, z_{i+1} = z_i v_i: the grand products of the permutation / lookup arguments
inline DeviceFr running_product(const DeviceFr& v, bool inclusive = false) {
    DeviceFr out(v.size());
    check(b200zk_fr_running_product_dev(v.ptr(), out.ptr(), v.size(), nullptr, inclusive ? 1u : 0u, nullptr));
    return out;
}
inline DeviceFr batch_invert(const DeviceFr& v) {
    DeviceFr out(v.size());
    check(b200zk_fr_batch_invert_dev(v.ptr(), out.ptr(), v.size(), nullptr));
    return out;
}
}  // namespace poly

// The numerator of the quotient polynomial as a device-resident register program (b200zk.h, gate programs).
class GateProgram {
public:
    GateProgram(const std::vector<uint32_t>& words, const std::vector<Fr>& consts, const std::vector<int32_t>& rotations,
                uint32_t n_columns, uint32_t k, uint32_t extended_k, const std::vector<Fr>& t_inv = {})
        : ek_(extended_k), ncol_(n_columns) {
        uint32_t log_period = 0;
        while ((size_t(1) << log_period) < t_inv.size()) log_period++;
        check(b200zk_gate_program_create(words.data(), (uint32_t)(words.size() / 4), reinterpret_cast<const uint8_t*>(consts.data()),
                                         (uint32_t)consts.size(), rotations.data(), (uint32_t)rotations.size(),
                                         t_inv.empty() ? nullptr : reinterpret_cast<const uint8_t*>(t_inv.data()), log_period,
                                         n_columns, k, extended_k, &h_));
    }
    ~GateProgram() { if (h_) b200zk_gate_program_release(h_); }
    GateProgram(const GateProgram&) = delete;
    GateProgram& operator=(const GateProgram&) = delete;
    void set_challenge(uint32_t index, const Fr& v) { check(b200zk_gate_program_set_const(h_, index, v.data())); }
    // evaluates the program at every row of the extended domain
    DeviceFr run(const std::vector<const DeviceFr*>& columns) const {
        if (columns.size() != ncol_) throw Error(B200ZK_ERR_INVALID_ARG, "gate program: wrong number of columns");
        DeviceFr out(uint64_t(1) << ek_);
        std::vector<const void*> ptrs;
        for (auto* c : columns) ptrs.push_back(c->ptr());
        check(b200zk_gate_program_run_dev(h_, ptrs.data(), out.ptr(), 0, nullptr));
        return out;
    }

private:
    uint64_t h_ = 0;
    uint32_t ek_, ncol_;
};

// ---------------------------------------------------------------------------------------------------
// The composed flows (csrc/h2mo.cu): transcript, multi_open with resident polynomials, multi_prepare, guards.
// ---------------------------------------------------------------------------------------------------
using G1Compressed = std::array<uint8_t, 48>;

// CardanoFriendlyBlake2b (/root/reference/src/plutus_gen/adjusted_types/mod.rs:30-72)
class Transcript {
public:
    Transcript() { check(b200zk_transcript_new(&h_)); }
    ~Transcript() { if (h_) b200zk_transcript_free(h_); }
    Transcript(const Transcript&) = delete;
    Transcript& operator=(const Transcript&) = delete;
    void common_scalar(const Fr& s) { check(b200zk_transcript_common_scalar(h_, s.data())); }
    void common_point(const G1Compressed& p) { check(b200zk_transcript_common_point(h_, p.data())); }
    Fr squeeze_challenge() { Fr c{}; check(b200zk_transcript_squeeze(h_, c.data())); return c; }
    uint64_t handle() const { return h_; }

private:
    uint64_t h_ = 0;
};

struct ProverQuery { uint32_t poly; Fr point; };                 // open polynomial `poly` at `point`
struct VerifierQuery { uint32_t commitment; Fr point; Fr eval; }; // commitment is claimed to evaluate to `eval` at `point`

// KZGCommitmentScheme::multi_open (message order /root/reference/src/plutus_gen/extraction/pcs/kzg.rs:55-79): returns
// f || q_evals || pi and leaves the transcript where the verifier's will be
inline std::vector<uint8_t> multi_open(const ParamsKZG& params, Transcript& t, const std::vector<const DeviceFr*>& polys,
                                       const std::vector<ProverQuery>& queries) {
    std::vector<const void*> ptrs;
    for (auto* p : polys) ptrs.push_back(p->ptr());
    std::vector<uint32_t> qp;
    std::vector<uint8_t> pts;
    for (auto& q : queries) { qp.push_back(q.poly); pts.insert(pts.end(), q.point.begin(), q.point.end()); }
    std::vector<uint8_t> proof(48 + 32 * queries.size() + 48);
    size_t len = 0;
    check(b200zk_h2mo_open_dev(params.g(), t.handle(), ptrs.data(), (uint32_t)polys.size(), polys.empty() ? 0 : polys[0]->size(), qp.data(),
                               pts.data(), (uint32_t)queries.size(), proof.data(), proof.size(), &len));
    proof.resize(len);
    return proof;
}

// The verifier's guard (upstream DualMSM): accept iff e(left, [s]G2) == e(right, G2); the pairing stays with the caller.
class Guard {
public:
    explicit Guard(uint64_t h) : h_(h) {}
    ~Guard() { if (h_) b200zk_guard_free(h_); }
    Guard(Guard&& o) noexcept : h_(o.h_) { o.h_ = 0; }
    Guard(const Guard&) = delete;
    Guard& operator=(const Guard&) = delete;
    std::pair<G1Affine, G1Affine> eval() const {
        G1Affine l{}, r{};
        check(b200zk_guard_eval(&h_, 1, nullptr, l.data(), r.data()));
        return {l, r};
    }
    uint64_t handle() const { return h_; }

private:
    uint64_t h_;
};
// KZGCommitmentScheme::multi_prepare (/root/reference/aiken-verifier/aiken_halo2/lib/halo2_kzg.ak:15-171)
inline Guard multi_prepare(Transcript& t, const std::vector<G1Compressed>& commitments, const std::vector<VerifierQuery>& queries,
                           const std::vector<uint8_t>& proof) {
    std::vector<uint32_t> qc;
    std::vector<uint8_t> pts, evs;
    for (auto& q : queries) {
        qc.push_back(q.commitment);
        pts.insert(pts.end(), q.point.begin(), q.point.end());
        evs.insert(evs.end(), q.eval.begin(), q.eval.end());
    }
    uint64_t g = 0;
    check(b200zk_h2mo_prepare(t.handle(), reinterpret_cast<const uint8_t*>(commitments.data()), (uint32_t)commitments.size(), qc.data(),
                              pts.data(), evs.data(), (uint32_t)queries.size(), proof.data(), proof.size(), &g, nullptr));
    return Guard(g);
}
// batch_verify up to the pairing (/root/reference/src/circuits/schnorr_circuit.rs:224-229): sum_i challenges[i] * guards[i]
inline std::pair<G1Affine, G1Affine> batch_guards(const std::vector<const Guard*>& guards, const std::vector<Fr>& challenges) {
    std::vector<uint64_t> hs;
    for (auto* g : guards) hs.push_back(g->handle());
    G1Affine l{}, r{};
    check(b200zk_guard_eval(hs.data(), (uint32_t)hs.size(), reinterpret_cast<const uint8_t*>(challenges.data()), l.data(), r.data()));
    return {l, r};
}

// Decompresses the commitments a proof carries (48 bytes each); throws on a bad encoding like the in-tree
// uncompress (/root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/CompressUncompress.hs:70-100).
inline std::vector<G1Affine> g1_decompress_batch(const std::vector<std::array<uint8_t, 48>>& compressed) {
    std::vector<G1Affine> out(compressed.size());
    if (!compressed.empty())
        check(b200zk_g1_decompress_batch(reinterpret_cast<const uint8_t*>(compressed.data()), compressed.size(),
                                         reinterpret_cast<uint8_t*>(out.data()), nullptr));
    return out;
}

}  // namespace b200zk
