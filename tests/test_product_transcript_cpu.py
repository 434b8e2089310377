"""The library's own transcript and multi-open scalar pipeline (csrc/h2mo.cu: host-side code of the PRODUCT, no GPU
needed) against every known answer the reference's tests hold for them: the challenge KATs of transcript.ak, the
golden 1 120-byte simple_mul proof with its pinned challenges, and the q_eval_sets / f_eval / v values 'extracted from
rust version of multi open' (Halo2MultiOpenMSM.hs:26-42)."""
import ctypes as C


def H(x):
    return int(x, 16)


def test_transcript_challenge_kats(zk, kats, pyref):
    t = kats["transcript"]
    T = zk.host.Transcript
    tr = T()
    tr.common_scalar(H(t["repr_only"]["repr"]))
    # (the KAT squeezes on a transcript whose proof stream is one zero byte that is never read)
    assert tr.squeeze_challenge() == H(t["repr_only"]["challenge"])
    tr = T()
    tr.common_scalar(1); tr.common_scalar(42)
    assert tr.squeeze_challenge() == H(t["after_scalar_42"])
    tr = T()
    tr.common_scalar(1); tr.common_point(pyref.g1_compress(pyref.g1_mul(pyref.G1_GEN, 42)))
    assert tr.squeeze_challenge() == H(t["after_point_42G"])
    tr = T(bytes.fromhex(t["mixed"]["proof"]))
    tr.common_scalar(1); tr.common_scalar(42)
    assert tr.read_point() == pyref.g1_compress(pyref.g1_neg(pyref.G1_GEN))
    assert tr.read_scalar() == H(t["mixed"]["scalar"])
    assert tr.squeeze_challenge() == H(t["mixed"]["challenge"])
    # consecutive squeezes differ (the 0x00 marker stays in the history) and match the checker's
    ref = pyref.Transcript()
    tr = T()
    for k in range(5):
        ref.common_scalar(k * 7 + 1); tr.common_scalar(k * 7 + 1)
        assert tr.squeeze_challenge() == ref.squeeze()
        assert tr.squeeze_challenge() == ref.squeeze()


def test_golden_simple_mul_proof_replay(zk, kats):
    gp = kats["transcript"]["golden_proof"]
    proof = bytes.fromhex(gp["proof"])
    ex = {k: H(v) for k, v in gp["expected"].items()}
    tr = zk.host.Transcript(proof)
    tr.common_scalar(H(gp["repr"]))
    tr.common_scalar(len(gp["public_inputs"]))
    for pi in gp["public_inputs"]:
        tr.common_scalar(pi)
    tr.read_point(); tr.read_point()
    tr.squeeze_challenge(); tr.squeeze_challenge()
    assert tr.squeeze_challenge() == ex["gamma"]
    for _ in range(4):
        tr.read_point()
    assert tr.squeeze_challenge() == ex["y"]
    tr.read_point(); tr.read_point()
    assert tr.squeeze_challenge() == ex["x"]
    evals = [tr.read_scalar() for _ in range(17)]
    assert evals[:3] == [ex["adviceEval1"], ex["adviceEval2"], ex["adviceEval3"]]
    assert tr.squeeze_challenge() == ex["x1"] and tr.squeeze_challenge() == ex["x2"]
    tr.read_point()
    assert tr.squeeze_challenge() == ex["x3"]
    for _ in range(3):
        tr.read_scalar()
    assert tr.squeeze_challenge() == ex["x4"]
    assert tr.read_point().hex() == gp["pi_compressed"] and tr.pos == 1120


def test_h2mo_scalar_pipeline_kat(zk, kats):
    """b200zk_h2mo_scalars on the ProofData.hs fixture: same q_eval_sets (matched point by point), f_eval and v."""
    h = kats["h2mo"]
    S = {k: H(v) for k, v in h["scalars"].items()}
    fr = zk.host.fr_bytes
    cmap = h["commitment_map"]
    queries = [(ci, S[p], S[e]) for ci, c in enumerate(cmap) for p, e in zip(c["points"], c["evals"])]
    qc = (C.c_uint32 * len(queries))(*[q[0] for q in queries])
    pts = b"".join(fr(q[1]) for q in queries)
    evs = b"".join(fr(q[2]) for q in queries)
    ch = b"".join(fr(S[x]) for x in ("x1", "x2", "x3", "x4"))
    pq = b"".join(fr(S[n]) for n in h["proof_x3_q_evals"])
    n_sets = len(h["point_sets"])
    total = sum(len(ps) for ps in h["point_sets"])
    qes = C.create_string_buffer(32 * total)
    f_eval, v = C.create_string_buffer(32), C.create_string_buffer(32)
    zk.capi.check(zk.lib().b200zk_h2mo_scalars(len(cmap), C.addressof(qc), zk.capi.addr(pts), zk.capi.addr(evs), len(queries),
                                               zk.capi.addr(ch), zk.capi.addr(pq), n_sets, zk.capi.addr(qes), len(qes),
                                               zk.capi.addr(f_eval), zk.capi.addr(v)))
    assert int.from_bytes(f_eval.raw, "little") == H(h["expected_f_eval"])
    assert int.from_bytes(v.raw, "little") == H(h["expected_v"])
    # the fixture lists the points of a set in rotation order, the library in ascending canonical order: match by point
    off = 0
    for ps, want in zip(h["point_sets"], h["expected_q_eval_sets"]):
        by_point = dict(zip((S[p] for p in ps), (H(x) for x in want)))
        for x in sorted(by_point):
            assert int.from_bytes(qes.raw[off:off + 32], "little") == by_point[x]
            off += 32
    # set structure: the fixture's own set indices are the numbering by first appearance that the library uses
    seen = {}
    for c in cmap:
        key = tuple(sorted(S[p] for p in c["points"]))
        assert seen.setdefault(key, len(seen)) == c["set"]


def test_h2mo_argument_errors(zk):
    rc = zk.lib().b200zk_transcript_squeeze(987654321, zk.capi.addr(C.create_string_buffer(32)))
    assert rc == -4
    out = C.create_string_buffer(96)
    assert zk.lib().b200zk_guard_eval(None, 0, None, zk.capi.addr(out), zk.capi.addr(out)) == -1


def test_h2mo_scalars_random_structures_vs_restatement(zk, pyref):
    """The library's host field arithmetic (64-bit-limb Montgomery products in csrc/h2mo.cu) against Python integers on random
    query structures whose points, evaluations and challenges include the edges of the field (0, 1, r - 1, 2^k neighbours)."""
    import random
    R = pyref.R_MOD
    rnd = random.Random(2024)
    edge = [0, 1, 2, R - 1, R - 2, (1 << 64) - 1, 1 << 64, (1 << 128) - 1, 1 << 192, (1 << 254) % R, R >> 1, (R >> 1) + 1]
    fr = zk.host.fr_bytes

    def pick():
        return rnd.choice(edge) if rnd.random() < 0.3 else rnd.randrange(R)

    for trial in range(40):
        n_comm = rnd.randrange(1, 9)
        n_pts = rnd.randrange(1, 5)
        pts_pool = []
        while len(pts_pool) < n_pts:                      # distinct points; x3 must differ from all of them
            p = pick()
            if p not in pts_pool:
                pts_pool.append(p)
        queries = []
        for ci in range(n_comm):
            for p in rnd.sample(pts_pool, rnd.randrange(1, n_pts + 1)):
                queries.append((ci, p, pick()))
        rnd.shuffle(queries)
        x1, x2, x4 = pick(), pick(), pick()
        x3 = pick()
        while x3 in pts_pool:
            x3 = rnd.randrange(R)
        psets, members, ev = pyref.h2mo_sets(queries)
        pq = [pick() for _ in psets]
        cmap = [(None, si, psets[si], ev[p]) for si, mem in enumerate(members) for p in mem]
        want_qes = pyref.h2mo_q_eval_sets(cmap, len(psets), x1)
        want_f = pyref.h2mo_f_eval(psets, want_qes, x2, x3, pq)
        want_v = pyref.h2mo_v(want_f, x4, pq)
        qc = (C.c_uint32 * len(queries))(*[q[0] for q in queries])
        pts = b"".join(fr(q[1]) for q in queries)
        evs = b"".join(fr(q[2]) for q in queries)
        ch = b"".join(fr(x) for x in (x1, x2, x3, x4))
        pqb = b"".join(fr(x) for x in pq)
        total = sum(len(ps) for ps in psets)
        qes = C.create_string_buffer(32 * total)
        f_eval, v = C.create_string_buffer(32), C.create_string_buffer(32)
        zk.capi.check(zk.lib().b200zk_h2mo_scalars(n_comm, C.addressof(qc), zk.capi.addr(pts), zk.capi.addr(evs), len(queries),
                                                   zk.capi.addr(ch), zk.capi.addr(pqb), len(psets), zk.capi.addr(qes), len(qes),
                                                   zk.capi.addr(f_eval), zk.capi.addr(v)))
        got = [int.from_bytes(qes.raw[32 * i:32 * i + 32], "little") for i in range(total)]
        flat = [want_qes[si][j] for si, ps in enumerate(psets) for j in range(len(ps))]
        assert got == flat, trial
        assert int.from_bytes(f_eval.raw, "little") == want_f, trial
        assert int.from_bytes(v.raw, "little") == want_v, trial
