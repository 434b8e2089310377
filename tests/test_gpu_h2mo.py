"""The composed flows (csrc/h2mo.cu): multi_open on resident polynomials, multi_prepare, guards and batched guards.
Checked three ways: (1) the proof bytes equal the big-integer restatement of the protocol (oracle/pyref.h2mo_open, whose
scalar side is pinned to the reference's known answers); (2) the verifier accepts: with the SRS secret s known to the
test, e(left, [s]G2) == e(right, G2) is the G1 identity s * left == right; (3) any tampering breaks it."""
import ctypes as C
import random

import pytest

pytestmark = pytest.mark.gpu

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def fr(x):
    return (x % R).to_bytes(32, "little")


def setup(gpu, pyref, k, n_polys, seed, s):
    rnd = random.Random(seed)
    n = 1 << k
    params = gpu.host.params_unsafe_setup(k, s)
    polys = [[rnd.randrange(R) for _ in range(n)] for _ in range(n_polys)]
    vecs = [gpu.host.FrVec.from_ints(p) for p in polys]
    w = pyref.omega(k)
    x = rnd.randrange(R)
    rot = {"cur": x, "next": x * w % R, "prev": x * pyref.fr_inv(w) % R, "last": x * pow(w, n - 7, R) % R}
    shapes = [["cur"], ["cur", "next"], ["cur", "next", "last"], ["cur", "prev"], ["cur", "next"]]
    queries = []
    for i in range(n_polys):
        for r in shapes[i % len(shapes)]:
            queries.append((i, rot[r]))
    # interleave them the way a prover's query list does (by rotation, not by polynomial)
    rnd.shuffle(queries)
    return params, polys, vecs, queries


def commitments_of(gpu, params, polys):
    pts = [gpu.host.KZGCommitmentScheme.commit(params, b"".join(fr(c) for c in p)) for p in polys]
    return pts, [gpu.host.g1_compress(p) for p in pts]


def prefix(t, comps, evals):
    t.common_scalar(0x1234)
    for c in comps:
        t.common_point(c)
    for e in evals:
        t.common_scalar(e)


def test_multi_open_matches_restatement_and_verifies(gpu, oracle, pyref):
    k, s = 8, 0x5EC2E7 * 0x10001 + 12345
    params, polys, vecs, queries = setup(gpu, pyref, k, 7, 11, s)
    n = 1 << k
    aff, comps = commitments_of(gpu, params, polys)
    g_table = C.create_string_buffer(96 * n)
    gpu.capi.check(gpu.lib().b200zk_bases_read(params._g, 0, n, gpu.capi.addr(g_table)))
    # commitments themselves: p(s) * G
    for p, a in zip(polys, aff):
        assert a == oracle.g1_mul(oracle.g1_generator(), fr(pyref.poly_eval(p, s)))
    evals = [pyref.poly_eval(polys[i], x) for i, x in queries]

    t_prover = gpu.host.Transcript()
    prefix(t_prover, comps, evals)
    proof = gpu.host.multi_open(params, t_prover, vecs, queries)

    # (1) bit-exact against the big-integer restatement driven by the checker's transcript
    t_ref = pyref.Transcript()
    t_ref.common_scalar(0x1234)
    for c in comps:
        t_ref.common_point_bytes(c)
    for e in evals:
        t_ref.common_scalar(e)

    def commit(coeffs):
        cb = b"".join(fr(c) for c in coeffs) + bytes(32 * (n - len(coeffs)))
        return pyref.g1_from_wire(oracle.msm(g_table.raw, cb, n))

    assert proof == pyref.h2mo_open(polys, queries, t_ref, commit)

    # (2) the verifier accepts
    t_ver = gpu.host.Transcript()
    prefix(t_ver, comps, evals)
    guard = gpu.host.multi_prepare(t_ver, comps, [(i, x, e) for (i, x), e in zip(queries, evals)], proof)
    left, right = guard.eval()
    assert left == oracle.g1_decompress(proof[-48:])[1]
    assert oracle.g1_mul(left, fr(s)) == right
    assert t_ver.squeeze_challenge() == t_prover.squeeze_challenge() == t_ref.squeeze()
    # the scalar side against the checker's pipeline (pinned to Halo2MultiOpenMSM.hs:26-42 in the CPU suite)
    sc = guard.scalars
    psets, members, ev = pyref.h2mo_sets([(i, x, e) for (i, x), e in zip(queries, evals)])
    cmap = [(None, si, psets[si], ev[p]) for si, mem in enumerate(members) for p in mem]
    qes = pyref.h2mo_q_eval_sets(cmap, len(psets), sc["x1"])
    pq = [int.from_bytes(proof[48 + 32 * i:80 + 32 * i], "little") for i in range(len(psets))]
    f_eval = pyref.h2mo_f_eval(psets, qes, sc["x2"], sc["x3"], pq)
    assert f_eval == sc["f_eval"] and pyref.h2mo_v(f_eval, sc["x4"], pq) == sc["v"]

    # (3) tampering: a wrong evaluation, a flipped proof scalar, a different commitment
    def accepts(comps_, qs_, proof_):
        t = gpu.host.Transcript()
        prefix(t, comps, evals)
        g = gpu.host.multi_prepare(t, comps_, qs_, proof_)
        l, r = g.eval()
        g.free()
        return oracle.g1_mul(l, fr(s)) == r

    good_q = [(i, x, e) for (i, x), e in zip(queries, evals)]
    assert accepts(comps, good_q, proof)
    bad_q = list(good_q)
    bad_q[3] = (bad_q[3][0], bad_q[3][1], (bad_q[3][2] + 1) % R)
    assert not accepts(comps, bad_q, proof)
    bad_proof = bytearray(proof)
    bad_proof[48] ^= 1
    assert not accepts(comps, good_q, bytes(bad_proof))
    swapped = [comps[1], comps[0]] + comps[2:]
    assert not accepts(swapped, good_q, proof)
    # a non-canonical scalar in the proof is rejected outright
    big = bytearray(proof)
    big[48:80] = (R + 5).to_bytes(32, "little")
    t = gpu.host.Transcript()
    prefix(t, comps, evals)
    with pytest.raises(gpu.B200zkError):
        gpu.host.multi_prepare(t, comps, good_q, bytes(big))
    guard.free()
    params.release()


def test_batched_guards_1024_proofs(gpu, oracle, pyref):
    """config 5: a batch of 1 024 opening proofs (64 distinct proofs of 24 commitments each, every one under its own random
    batching challenge) verified as ONE pair of sums: all 27 000 points decompressed on the GPU, two MSMs; accepted iff
    s * left == right.  One bad proof anywhere makes the batch fail."""
    k, s, n_polys = 6, 0xABCDEF0123456789, 24
    rnd = random.Random(5)
    guards = []
    params = None
    for j in range(64):
        params_j, polys, vecs, queries = setup(gpu, pyref, k, n_polys, 100 + j, s)
        if params is None:
            params = params_j
        else:
            params_j.release()
        _, comps = commitments_of(gpu, params, polys)
        evals = [pyref.poly_eval(polys[i], x) for i, x in queries]
        t = gpu.host.Transcript()
        prefix(t, comps, evals)
        proof = gpu.host.multi_open(params, t, vecs, queries)
        t = gpu.host.Transcript()
        prefix(t, comps, evals)
        guards.append(gpu.host.multi_prepare(t, comps, [(i, x, e) for (i, x), e in zip(queries, evals)], proof))
        for v in vecs:
            v.free()
    batch = [guards[i % 64] for i in range(1024)]
    ch = [rnd.randrange(1, R) for _ in range(1024)]
    left, right = gpu.host.batch_guards(batch, ch)
    assert oracle.g1_mul(left, fr(s)) == right
    # left is sum c_i pi_i: against the checker on the decompressed pi points
    # one bad guard: prepared against a wrong evaluation
    params_b, polys, vecs, queries = setup(gpu, pyref, k, n_polys, 999, s)
    _, comps = commitments_of(gpu, params, polys)
    evals = [pyref.poly_eval(polys[i], x) for i, x in queries]
    t = gpu.host.Transcript()
    prefix(t, comps, evals)
    proof = gpu.host.multi_open(params, t, vecs, queries)
    t = gpu.host.Transcript()
    prefix(t, comps, evals)
    bad = gpu.host.multi_prepare(t, comps, [(i, x, (e + (1 if q == 0 else 0)) % R) for q, ((i, x), e) in enumerate(zip(queries, evals))], proof)
    batch[517] = bad
    left, right = gpu.host.batch_guards(batch, ch)
    assert oracle.g1_mul(left, fr(s)) != right
    params_b.release()
    params.release()


def test_multi_open_concurrent_callers_and_changing_sizes(gpu, oracle, pyref):
    """The opening flow keeps its scratch in per-call arenas that are reused between calls: two host threads open different
    polynomial sets at the same time, then the sizes change (k = 9, 6, 10 -- an arena shrinks in use and grows again); every
    proof equals the one produced alone and verifies."""
    import threading
    s = 0x1D0C5 * 0x7777 + 99
    cases = {}
    for k in (9, 6, 10):
        params, polys, vecs, queries = setup(gpu, pyref, k, 5, 40 + k, s)
        _, comps = commitments_of(gpu, params, polys)
        evals = [pyref.poly_eval(polys[i], x) for i, x in queries]
        cases[k] = (params, vecs, queries, comps, evals)

    def open_one(k):
        params, vecs, queries, comps, evals = cases[k]
        t = gpu.host.Transcript()
        prefix(t, comps, evals)
        return gpu.host.multi_open(params, t, vecs, queries)

    alone = {k: open_one(k) for k in (9, 6, 10)}
    for k in (9, 6, 10):                                   # sizes in changing order, each reproducible
        assert open_one(k) == alone[k]
    got, errs = {}, []

    def worker(k, reps):
        try:
            for _ in range(reps):
                got.setdefault(k, []).append(open_one(k))
        except Exception as e:                             # pragma: no cover - reported below
            errs.append(e)

    th = [threading.Thread(target=worker, args=(9, 6)), threading.Thread(target=worker, args=(10, 6)),
          threading.Thread(target=worker, args=(6, 6))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for k in (9, 6, 10):
        assert all(p == alone[k] for p in got[k]) and len(got[k]) == 6, k
    # and the verifier accepts one of them
    params, vecs, queries, comps, evals = cases[10]
    t = gpu.host.Transcript()
    prefix(t, comps, evals)
    guard = gpu.host.multi_prepare(t, comps, [(i, x, e) for (i, x), e in zip(queries, evals)], alone[10])
    left, right = guard.eval()
    assert oracle.g1_mul(left, fr(s)) == right
    guard.free()
    for params, vecs, *_ in cases.values():
        for v in vecs:
            v.free()
        params.release()
