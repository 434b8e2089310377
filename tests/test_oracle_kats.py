"""Pins both oracles (oracle/pyref.py, oracle/b200zk_oracle.c) against every known-answer vector
the reference's own tests hold for this path (tests/golden/reference_kats.json, extracted by
tests/golden/extract_reference_kats.py from the Aiken / Plinth test files)."""
import random


def H(x):
    return int(x, 16)


def test_constants(pyref, kats):
    c = kats["constants"]
    assert pyref.R_MOD == H(c["r"]) and pyref.P_MOD == H(c["p"]) and pyref.DELTA == H(c["delta"])
    assert pyref.Transcript.R256 == H(kats["transcript"]["R256"])


def test_omega_convention(pyref, kats):
    o = kats["omega_k14"]
    w = pyref.omega(14)
    assert w == H(o["omega"]) and pyref.fr_inv(w) == H(o["omega_inv"])
    assert pow(w, 1 << 14, pyref.R_MOD) == 1 and pow(w, 1 << 13, pyref.R_MOD) == pyref.R_MOD - 1
    assert [pow(w, e, pyref.R_MOD) for e in range(-6, 1)] == [H(x) for x in o["rotations_m6_to_0"]]


def test_lagrange_basis(pyref, kats):
    l = kats["lagrange_k14"]
    assert H(l["barycentric_weight"]) == pyref.fr_inv(1 << 14)
    assert pow(H(l["x"]), 1 << 14, pyref.R_MOD) == H(l["xn"])
    got = pyref.lagrange_basis(H(l["x"]), H(l["xn"]), H(l["barycentric_weight"]), [H(x) for x in l["rotations"]])
    assert got == [H(x) for x in l["expected"]]


def test_g1_encoding_pyref(pyref, kats):
    e = kats["g1_encoding"]
    G = pyref.G1_GEN
    assert pyref.g1_compress(G).hex() == e["generator"]
    assert pyref.g1_compress(pyref.g1_neg(G)).hex() == e["neg_generator"]
    assert pyref.g1_compress(pyref.g1_mul(G, 42)).hex() == e["g_times_42"]
    assert pyref.g1_decompress(bytes.fromhex(e["g_times_42"])) == pyref.g1_mul(G, 42)
    assert pyref.fr_from_le(bytes.fromhex(e["scalar_r_bytes"])) == 0
    assert pyref.fr_from_le(bytes.fromhex(e["scalar_overflow_bytes"])) == H(e["scalar_overflow_value"])


def test_g1_encoding_c_oracle(oracle, pyref, kats):
    e = kats["g1_encoding"]
    g = oracle.g1_generator()
    assert pyref.g1_from_wire(g) == pyref.G1_GEN
    assert oracle.g1_compress(g).hex() == e["generator"]
    p42 = oracle.g1_mul(g, pyref.fr_to_le(42))
    assert oracle.g1_compress(p42).hex() == e["g_times_42"]
    rc, dec = oracle.g1_decompress(bytes.fromhex(e["neg_generator"]))
    assert rc == 0 and pyref.g1_from_wire(dec) == pyref.g1_neg(pyref.G1_GEN)
    # reduce-on-read of the scalar wire format: r -> 0, over-range -> value - r
    assert oracle.g1_mul(g, bytes.fromhex(e["scalar_r_bytes"])) == bytes(96)
    over = oracle.g1_mul(g, bytes.fromhex(e["scalar_overflow_bytes"]))
    assert over == oracle.g1_mul(g, pyref.fr_to_le(H(e["scalar_overflow_value"])))


def test_transcript_challenges(pyref, kats):
    t = kats["transcript"]
    tr = pyref.Transcript(b"\x00")
    tr.common_scalar(H(t["repr_only"]["repr"]))
    assert tr.squeeze() == H(t["repr_only"]["challenge"])
    tr = pyref.Transcript()
    tr.common_scalar(1); tr.common_scalar(42)
    assert tr.squeeze() == H(t["after_scalar_42"])
    tr = pyref.Transcript()
    tr.common_scalar(1); tr.common_point(pyref.g1_mul(pyref.G1_GEN, 42))
    assert tr.squeeze() == H(t["after_point_42G"])
    tr = pyref.Transcript(bytes.fromhex(t["mixed"]["proof"]))
    tr.common_scalar(1); tr.common_scalar(42)
    assert tr.read_point() == pyref.g1_neg(pyref.G1_GEN)
    assert tr.read_scalar() == H(t["mixed"]["scalar"])
    assert tr.squeeze() == H(t["mixed"]["challenge"])


def test_golden_simple_mul_proof(pyref, oracle, kats):
    """Replays the 1 120-byte simple_mul proof: layout 8 G1 | 17 Fr | f | 3 Fr | pi, every point
    decompresses onto the curve in both oracles, every pinned challenge reproduces."""
    gp = kats["transcript"]["golden_proof"]
    proof = bytes.fromhex(gp["proof"])
    assert len(proof) == 8 * 48 + 17 * 32 + 48 + 3 * 32 + 48 == 1120
    ex = {k: H(v) for k, v in gp["expected"].items()}
    tr = pyref.Transcript(proof)
    tr.common_scalar(H(gp["repr"]))
    tr.common_scalar(len(gp["public_inputs"]))
    for pi in gp["public_inputs"]:
        tr.common_scalar(pi)
    pts = [tr.read_point(), tr.read_point()]
    tr.squeeze(); tr.squeeze()
    assert tr.squeeze() == ex["gamma"]
    pts += [tr.read_point() for _ in range(4)]
    assert tr.squeeze() == ex["y"]
    pts += [tr.read_point() for _ in range(2)]
    assert tr.squeeze() == ex["x"]
    evals = [tr.read_scalar() for _ in range(17)]
    assert evals[:3] == [ex["adviceEval1"], ex["adviceEval2"], ex["adviceEval3"]]
    assert tr.squeeze() == ex["x1"] and tr.squeeze() == ex["x2"]
    pts.append(tr.read_point())
    assert tr.squeeze() == ex["x3"]
    [tr.read_scalar() for _ in range(3)]
    assert tr.squeeze() == ex["x4"]
    pi_bytes = tr.read_point_bytes()
    assert pi_bytes.hex() == gp["pi_compressed"] and tr.pos == 1120
    pts.append(pyref.g1_decompress(pi_bytes))
    for p in pts:
        assert pyref.g1_is_on_curve(p)
        w = pyref.g1_to_wire(p)
        c = pyref.g1_compress(p)
        assert oracle.g1_compress(w) == c and oracle.g1_decompress(c) == (0, w)
        assert oracle.L.orc_g1_on_curve(w) == 1


def test_h2mo_scalar_pipeline(pyref, kats):
    """q_eval_sets, f_eval and v of the verifier's multi-open (values 'extracted from rust version
    of multi open', Halo2MultiOpenMSM.hs:24-42)."""
    h = kats["h2mo"]
    S = {k: H(v) for k, v in h["scalars"].items()}
    cmap = [(c["commitment"], c["set"], [S[p] for p in c["points"]], [S[e] for e in c["evals"]]) for c in h["commitment_map"]]
    q = pyref.h2mo_q_eval_sets(cmap, 3, S["x1"])
    assert q == [[H(x) for x in s] for s in h["expected_q_eval_sets"]]
    psets = [[S[p] for p in ps] for ps in h["point_sets"]]
    pq = [S[n] for n in h["proof_x3_q_evals"]]
    f_eval = pyref.h2mo_f_eval(psets, q, S["x2"], S["x3"], pq)
    assert f_eval == H(h["expected_f_eval"])
    assert pyref.h2mo_v(f_eval, S["x4"], pq) == H(h["expected_v"])
    for name, (x, y) in h["points"].items():
        assert pyref.g1_is_on_curve((H(x), H(y))), name


def test_c_oracle_field_vs_python(oracle, pyref):
    rnd = random.Random(1)
    for _ in range(300):
        a, b = rnd.randrange(pyref.R_MOD), rnd.randrange(pyref.R_MOD)
        le = lambda v, n: v.to_bytes(n, "little")
        assert oracle.field("orc_fr_mul", le(a, 32), le(b, 32)) == le(a * b % pyref.R_MOD, 32)
        assert oracle.field("orc_fr_sub", le(a, 32), le(b, 32)) == le((a - b) % pyref.R_MOD, 32)
        a, b = rnd.randrange(pyref.P_MOD), rnd.randrange(pyref.P_MOD)
        assert oracle.field("orc_fp_mul", le(a, 48), le(b, 48)) == le(a * b % pyref.P_MOD, 48)
        assert oracle.field("orc_fp_add", le(a, 48), le(b, 48)) == le((a + b) % pyref.P_MOD, 48)
    a = rnd.randrange(1, pyref.P_MOD)
    assert oracle.field("orc_fp_inv", a.to_bytes(48, "little")) == pow(a, -1, pyref.P_MOD).to_bytes(48, "little")


def test_c_oracle_msm_cross_checks(oracle, pyref):
    """double-and-add (pyref) == double-and-add (C) == Pippenger (C), incl. edge cases."""
    rnd = random.Random(2)
    n = 150
    pts = [pyref.g1_mul(pyref.G1_GEN, rnd.randrange(1, pyref.R_MOD)) for _ in range(n)]
    pts[3] = pyref.INF
    pts[5] = pts[4]
    pts[7] = pyref.g1_neg(pts[6])
    sc = [rnd.randrange(pyref.R_MOD) for _ in range(n)]
    sc[0], sc[1], sc[2], sc[6], sc[7] = 0, 1, pyref.R_MOD - 1, 5, 5
    pb = b"".join(pyref.g1_to_wire(p) for p in pts)
    sb = b"".join(pyref.fr_to_le(s) for s in sc)
    exp = pyref.g1_to_wire(pyref.g1_msm_naive(pts, sc))
    assert oracle.msm(pb, sb, n, naive=True) == exp
    assert oracle.msm(pb, sb, n) == exp
    for m in (1, 2, 3, 17):
        assert oracle.msm(pb, sb, m) == oracle.msm(pb, sb, m, naive=True)
    assert oracle.msm(pb, sb, 0) == bytes(96)


def test_c_oracle_msm_dlog_property(oracle, pyref):
    """MSM(s, a_i*G) == (sum s_i a_i)*G on the synthetic inputs used by the bench."""
    import ctypes as C
    n = 5000
    bases = oracle.synth_bases(0xB200, 0, n)
    sc = oracle.synth_scalars(1, 0, n)
    a = [pyref.synth_base_dlog(0xB200, i) for i in range(n)]
    assert pyref.g1_from_wire(bases[:96]) == pyref.g1_mul(pyref.G1_GEN, a[0])
    dot = sum(int.from_bytes(sc[32 * i:32 * i + 32], "little") * a[i] for i in range(n)) % pyref.R_MOD
    arr = (C.c_uint64 * n)(*a)
    out = C.create_string_buffer(32)
    oracle.L.orc_fr_dot_u64(sc, arr, n, out)
    assert int.from_bytes(out.raw, "little") == dot
    assert oracle.msm(bases, sc, n) == oracle.g1_mul(oracle.g1_generator(), pyref.fr_to_le(dot))


def test_c_oracle_ntt_cross_checks(oracle, pyref):
    rnd = random.Random(3)
    for k in range(0, 9):
        a = [rnd.randrange(pyref.R_MOD) for _ in range(1 << k)]
        data = b"".join(pyref.fr_to_le(x) for x in a)
        w = pyref.omega(k)
        exp = b"".join(pyref.fr_to_le(x) for x in pyref.ntt_naive(a, w))
        assert oracle.ntt(data, k, pyref.fr_to_le(w)) == exp
        assert oracle.ntt_naive(data, k, pyref.fr_to_le(w)) == exp
        assert oracle.ntt(exp, k, pyref.fr_to_le(pyref.fr_inv(w)), 1) == data
        cos = b"".join(pyref.fr_to_le(x) for x in pyref.coset_ntt(a, w, 7))
        assert oracle.ntt(data, k, pyref.fr_to_le(w), 0, pyref.fr_to_le(7)) == cos
        assert oracle.ntt(cos, k, pyref.fr_to_le(pyref.fr_inv(w)), 1, None, pyref.fr_to_le(pyref.fr_inv(7))) == data


def test_c_oracle_synthetic_bases_windowed_equals_double_and_add(oracle, pyref):
    """The fast base generator of the CPU arm (fixed-base windows + one inversion per block of 256) against the plain
    double-and-add definition, across block boundaries and for a start offset, and against the big-integer definition."""
    import ctypes as C
    L = oracle.L
    L.orc_g1_synth_bases_naive.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_char_p, C.c_int]
    for start, n in ((0, 1), (0, 255), (0, 256), (5, 257), (1000, 1200)):
        a = C.create_string_buffer(96 * n)
        L.orc_g1_synth_bases_naive(0xB200, start, n, a, 0)
        assert oracle.synth_bases(0xB200, start, n) == a.raw, (start, n)
    assert pyref.g1_from_wire(oracle.synth_bases(0xB200, 3, 1)) == pyref.g1_mul(pyref.G1_GEN, pyref.synth_base_dlog(0xB200, 3))
