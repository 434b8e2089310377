"""Generates tests/golden/oracle_vectors.json: small fixed input/output vectors for the MSM, the NTT and the
polynomial-side rows, computed by the big-int oracle (oracle/pyref.py) and cross-checked against the C oracle where
it has the operation.  The reference's own MSM / NTT implementation cannot run in this environment (Rust crates that are
not vendored), so these are ORACLE-generated vectors: they pin the agreement of pyref, the C oracle and the CUDA path
on identical inputs and guard all three against drift.  Inputs are the seeded synthetic generators (splitmix64).
usage: python tests/golden/make_oracle_vectors.py   (rewrites the JSON next to it)"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyref as P  # noqa: E402
from conftest import Oracle, _build_oracle  # noqa: E402


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def fr_vec(seed, n):
    return [P.splitmix64(seed * 1000003 + i) * P.splitmix64(seed + 17 * i + 5) % P.R_MOD for i in range(n)]


def to_bytes(v):
    return b"".join(x.to_bytes(32, "little") for x in v)


def main():
    orc = Oracle(_build_oracle())
    out = {"about": "oracle-generated vectors (pyref + C oracle agree); inputs: synth_bases(0xB200, 0, n), synth_scalars(seed, 0, n), "
                    "fr_vec(seed, n)[i] = splitmix64(seed*1000003+i) * splitmix64(seed+17i+5) mod r"}
    # --- MSM: C oracle Pippenger == C oracle naive == pyref double-and-add
    out["msm"] = []
    for n, seed in ((1, 3), (7, 4), (64, 5), (1000, 6)):
        bases, sc = orc.synth_bases(0xB200, 0, n), orc.synth_scalars(seed, 0, n)
        r = orc.msm(bases, sc, n)
        assert r == orc.msm(bases, sc, n, naive=True)
        if n <= 64:
            pts = [P.g1_from_wire(bases[96 * i:96 * i + 96]) for i in range(n)]
            ks = [int.from_bytes(sc[32 * i:32 * i + 32], "little") for i in range(n)]
            assert P.g1_to_wire(P.g1_msm_naive(pts, ks)) == r
        out["msm"].append({"n": n, "scalar_seed": seed, "result_compressed": orc.g1_compress(r).hex()})
    # --- NTT: C oracle radix-2 == pyref radix-2 == O(n^2) DFT (small)
    out["ntt"] = []
    for log_n in (3, 8, 12):
        n = 1 << log_n
        v = fr_vec(20 + log_n, n)
        w = P.omega(log_n)
        r = P.ntt(v, w)
        assert to_bytes(r) == orc.ntt(to_bytes(v), log_n, w.to_bytes(32, "little"))
        if log_n <= 8:
            assert r == P.ntt_naive(v, w)
        cos = P.coset_ntt(v, w, 7)
        out["ntt"].append({"log_n": log_n, "seed": 20 + log_n, "forward_sha256": sha(to_bytes(r)), "first": hex(r[0]), "last": hex(r[-1]),
                           "coset7_forward_sha256": sha(to_bytes(cos))})
    # --- polynomial side
    c = fr_vec(31, 300)
    z = fr_vec(32, 1)[0]
    q, e = P.kate_div(c, z)
    out["kate_div"] = {"n": 300, "seed": 31, "z": hex(z), "eval": hex(e), "quotient_sha256": sha(to_bytes(q))}
    v = fr_vec(33, 500)
    out["running_product"] = {"n": 500, "seed": 33, "exclusive_sha256": sha(to_bytes(P.running_product(v))),
                              "last_inclusive": hex(P.running_product(v, 1, True)[-1])}
    vz = list(v)
    vz[7] = vz[100] = 0
    out["batch_invert"] = {"n": 500, "seed": 33, "zeros_at": [7, 100], "sha256": sha(to_bytes(P.batch_invert(vz)))}
    # gate program: the arithmetic gate q_m a b + q_l a + q_r b - c with a rotated term, over 5 columns at k = 4, ext_k = 6
    k, ext_k = 4, 6
    cols = [fr_vec(40 + i, 1 << ext_k) for i in range(5)]
    prog = [(P.GATE_OPS["mul"], 1, P.gate_col(0, 0), P.gate_col(1, 0), 0),
            (P.GATE_OPS["mul"], 1, P.gate_reg(1), P.gate_col(3, 0), 0),
            (P.GATE_OPS["muladd"], 1, P.gate_col(4, 0), P.gate_col(0, 1), P.gate_reg(1)),
            (P.GATE_OPS["sub"], 0, P.gate_reg(1), P.gate_col(2, 2), 0),
            (P.GATE_OPS["muladd"], 0, P.gate_reg(0), P.gate_const(0), P.gate_const(1))]
    consts = [fr_vec(50, 1)[0], 9]
    t_inv = P.vanishing_inverse_on_coset(7, k, ext_k)
    g = P.gate_eval(prog, consts, [0, 1, -1], cols, k, ext_k, t_inv)
    out["gate_program"] = {"k": k, "ext_k": ext_k, "column_seeds": [40, 41, 42, 43, 44], "const_seed": 50, "rotations": [0, 1, -1],
                           "words": P.gate_program_words(prog), "sha256": sha(to_bytes(g)), "row0": hex(g[0])}
    # SRS: both tables at k = 3 for a fixed secret
    s = fr_vec(60, 1)[0]
    mono, lag = P.srs_scalars(s, 3)
    G = orc.g1_generator()
    out["srs"] = {"k": 3, "secret": hex(s),
                  "g": [orc.g1_compress(orc.g1_mul(G, m.to_bytes(32, "little"))).hex() for m in mono],
                  "g_lagrange": [orc.g1_compress(orc.g1_mul(G, m.to_bytes(32, "little"))).hex() for m in lag]}
    for i, m in enumerate(mono[:3]):
        assert P.g1_compress(P.g1_mul(P.G1_GEN, m)).hex() == out["srs"]["g"][i]
    json.dump(out, open(os.path.join(HERE, "oracle_vectors.json"), "w"), indent=1)
    print("wrote oracle_vectors.json")


if __name__ == "__main__":
    main()
