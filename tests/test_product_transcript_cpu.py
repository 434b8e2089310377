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
