"""Batched verification up to the pairing (BASELINE configs[4]: batched verify_proof of 1 024 proofs; reference loop at
/root/reference/src/circuits/schnorr_circuit.rs:184-232): `distinct` real opening proofs over 24 commitments each are produced by
the library's own prover (b200zk_h2mo_open_dev) under an SRS whose secret the tool knows, prepared by b200zk_h2mo_prepare, and
1 024 guards (each distinct proof repeated under its own random batching challenge) are evaluated as ONE pair of sums by
b200zk_guard_eval: ~27 000 compressed points decompressed on the GPU, two ad-hoc MSMs (split over the bound GPUs).  Accepted iff
s * left == right, which the CPU checker verifies.

usage: python tools/verify_bench.py [--proofs 1024] [--distinct 16]"""
import argparse
import ctypes as C
import importlib
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def fr(x):
    return (x % R).to_bytes(32, "little")


def run_batched_verify(zk, L, n_proofs=1024, distinct=16, k=6, n_polys=24, reps=5):
    """L: the CPU checker (ctypes handle) or None (no accept check)."""
    H = zk.host
    rnd = random.Random(7)
    s = 0xABCDEF0123456789ABCDEF % R
    n = 1 << k
    params = H.params_unsafe_setup(k, s)
    w = pow(H.ROOT_OF_UNITY, 1 << (32 - k), R)
    guards, t_open, t_prep = [], 0.0, 0.0
    for j in range(distinct):
        polys = [[rnd.randrange(R) for _ in range(n)] for _ in range(n_polys)]
        vecs = [H.FrVec.from_ints(p) for p in polys]
        x = rnd.randrange(R)
        rot = [x, x * w % R, x * pow(w, R - 2, R) % R]
        queries = [(i, rot[0]) for i in range(n_polys)] + [(i, rot[1]) for i in range(0, n_polys, 3)] + [(i, rot[2]) for i in range(1, n_polys, 6)]
        comps = [H.g1_compress(c) for c in H.KZGCommitmentScheme.commit_batch(params, [b"".join(fr(c) for c in p) for p in polys])]
        evals = []
        for i, pt in queries:                      # Horner on the host: what create_proof wrote into the transcript earlier
            acc = 0
            for c in reversed(polys[i]):
                acc = (acc * pt + c) % R
            evals.append(acc)
        tr = H.Transcript()
        for c in comps:
            tr.common_point(c)
        for e in evals:
            tr.common_scalar(e)
        t0 = time.perf_counter()
        proof = H.multi_open(params, tr, vecs, queries)
        t_open += time.perf_counter() - t0
        tr = H.Transcript()
        for c in comps:
            tr.common_point(c)
        for e in evals:
            tr.common_scalar(e)
        t0 = time.perf_counter()
        guards.append(H.multi_prepare(tr, comps, [(i, pt, e) for (i, pt), e in zip(queries, evals)], proof))
        t_prep += time.perf_counter() - t0
        for v in vecs:
            v.free()
    batch = [guards[i % distinct] for i in range(n_proofs)]
    ch = [rnd.randrange(1, R) for _ in range(n_proofs)]
    left, right = H.batch_guards(batch, ch)
    t0 = time.perf_counter()
    for _ in range(reps):
        left, right = H.batch_guards(batch, ch)
    ms = (time.perf_counter() - t0) / reps * 1e3
    accepted = None
    if L is not None:
        out = C.create_string_buffer(96)
        L.orc_g1_mul.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.orc_g1_mul(left, fr(s), out)
        accepted = out.raw == right
    n_points = n_proofs * (n_polys + 3) + 1
    for g in guards:
        g.free()
    params.release()
    return {"proofs": n_proofs, "distinct_proofs": distinct, "commitments_per_proof": n_polys, "points_decompressed": n_points,
            "guard_eval_ms": ms, "proofs_per_s": n_proofs / (ms * 1e-3), "multi_open_ms_per_proof_k%d" % k: t_open / distinct * 1e3,
            "multi_prepare_ms_per_proof": t_prep / distinct * 1e3, "accepted_s_left_eq_right": accepted,
            "note": "one b200zk_guard_eval call: host scaling of the guards' terms by their batching challenges, one GPU decompression "
                    "batch, two ad-hoc MSMs (split over the bound GPUs), host clock; the pairing is the caller's"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--proofs", type=int, default=1024)
    ap.add_argument("--distinct", type=int, default=16)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    import bench
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    if args.gpus > 1:
        zk.capi.init_devices(None, args.gpus)
    else:
        zk.init(0)
    print(json.dumps(run_batched_verify(zk, bench.load_oracle(), args.proofs, args.distinct)))


if __name__ == "__main__":
    main()
