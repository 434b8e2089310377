"""SRS generation and batched decompression (row 4) against the oracle, through the C ABI."""
import random

import pytest

pytestmark = pytest.mark.gpu

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def fr(x):
    return (x % R).to_bytes(32, "little")


@pytest.mark.parametrize("k", [0, 1, 3, 6])
def test_srs_tables_vs_oracle(gpu, oracle, pyref, k):
    H = gpu.host
    s = 0x1234567890ABCDEF1234567890ABCDEF ** 2 % R
    n = 1 << k
    g, gl = H.srs_generate(s, k)
    mono, lag = pyref.srs_scalars(s, k)
    G = oracle.g1_generator()
    got_g, got_l = H.g1_export(g, n), H.g1_export(gl, n)
    for i in range(n):
        assert got_g[96 * i:96 * i + 96] == oracle.g1_mul(G, fr(mono[i])), ("g", i)
        assert got_l[96 * i:96 * i + 96] == oracle.g1_mul(G, fr(lag[i])), ("g_lagrange", i)
    assert got_g[:96] == G


def test_srs_commitments_agree_between_bases(gpu, oracle, pyref):
    """commit(coeffs) against g equals commit_lagrange(evals) against g_lagrange -- the identity the prover relies
    on -- and both equal p(s) * G."""
    H = gpu.host
    k = 10
    n = 1 << k
    s = random.Random(5).randrange(R)
    params = H.params_unsafe_setup(k, s)
    coeffs = oracle.synth_scalars(3, 0, n)
    dom = H.EvaluationDomain(4, k)
    evals = dom.coeff_to_lagrange(coeffs)
    c1 = H.KZGCommitmentScheme.commit(params, coeffs)
    c2 = H.KZGCommitmentScheme.commit_lagrange(params, evals)
    ps = pyref.poly_eval([int.from_bytes(coeffs[32 * i:32 * i + 32], "little") for i in range(n)], s)
    assert c1 == c2 == oracle.g1_mul(oracle.g1_generator(), fr(ps))
    params.release()


def test_srs_secret_in_domain_is_rejected(gpu, pyref):
    with pytest.raises(gpu.B200zkError):
        gpu.host.srs_generate(pyref.omega(4), 4)


def test_fixed_base_mul_edge_scalars(gpu, oracle):
    H = gpu.host
    import ctypes as C
    vals = [0, 1, 2, 255, 256, R - 1, R - 2, (1 << 248) - 1, 0xFF << 120]
    sc = H.FrVec(len(vals))
    sc.upload(b"".join(fr(v) for v in vals))          # canonical scalars
    out = H.DeviceBuffer(96 * len(vals))
    gpu.capi.check(gpu.lib().b200zk_g1_fixed_mul_dev(sc.ptr, gpu.FMT_CANONICAL, len(vals), out.ptr, None))
    got = H.g1_export(out, len(vals))
    G = oracle.g1_generator()
    for i, v in enumerate(vals):
        assert got[96 * i:96 * i + 96] == oracle.g1_mul(G, fr(v)), hex(v)


def test_decompress_batch_vs_oracle(gpu, oracle, kats, pyref):
    H = gpu.host
    n = 300
    pts = oracle.synth_bases(0xB200, 0, n)
    comp = b"".join(oracle.g1_compress(pts[96 * i:96 * i + 96]) for i in range(n))
    comp += bytes([0xC0]) + bytes(47)                                  # identity
    neg = pyref.g1_to_wire(pyref.g1_neg(pyref.g1_from_wire(pts[:96])))
    comp += oracle.g1_compress(neg)
    aff, status = H.g1_decompress_batch(comp)
    assert status == [0] * (n + 2)
    assert aff[:96 * n] == pts and aff[96 * n:96 * n + 96] == bytes(96) and aff[96 * n + 96:] == neg
    # the golden simple_mul proof of the reference's Aiken tests: all ten points decompress like the oracle's
    proof = bytes.fromhex(kats["transcript"]["golden_proof"]["proof"])
    offs = list(range(0, 384, 48)) + [384 + 17 * 32, 1120 - 48]
    blob = b"".join(proof[o:o + 48] for o in offs)
    aff, status = H.g1_decompress_batch(blob)
    assert status == [0] * len(offs)
    for j in range(len(offs)):
        rc, want = oracle.g1_decompress(blob[48 * j:48 * j + 48])
        assert rc == 0 and aff[96 * j:96 * j + 96] == want


def test_decompress_rejects_like_the_oracle(gpu, oracle):
    H = gpu.host
    good = oracle.g1_compress(oracle.g1_generator())
    bad = [
        bytes([good[0] & 0x7F]) + good[1:],                   # compression flag missing
        bytes([0xC0]) + bytes(46) + b"\x01",                  # infinity with a non-zero x
        bytes([0xE0]) + bytes(47),                            # infinity with the sign flag
        bytes([0x9F]) + b"\xff" * 47,                         # x >= p
    ]
    x = 1
    while True:                                               # an x that is not on the curve
        c = bytes([0x80]) + x.to_bytes(47, "big")
        if oracle.g1_decompress(c)[0] == -4:
            bad.append(c)
            break
        x += 1
    blob = good + b"".join(bad)
    aff, status = H.g1_decompress_batch(blob, strict=False)
    assert status == [0, 1, 2, 2, 3, 4]
    assert [-oracle.g1_decompress(b)[0] for b in bad] == status[1:]
    assert aff[96:] == bytes(96 * len(bad))
    with pytest.raises(gpu.B200zkError) as ei:
        H.g1_decompress_batch(blob)
    assert ei.value.code == -7


def test_batched_verifier_front_end(gpu, oracle):
    """Config 5's shape in miniature: decompress a batch of proof points on the GPU and feed them straight into the
    ad-hoc (DualMSM) sum; equals the oracle's MSM over the oracle's decompression."""
    H = gpu.host
    n = 2048
    pts = oracle.synth_bases(0xB200, 100, n)
    comp = b"".join(oracle.g1_compress(pts[96 * i:96 * i + 96]) for i in range(n))
    sc = oracle.synth_scalars(9, 0, n)
    aff, status = H.g1_decompress_batch(comp)
    assert not any(status)
    msm = H.DualMSM()
    for i in range(n):
        msm.append_right(int.from_bytes(sc[32 * i:32 * i + 32], "little"), aff[96 * i:96 * i + 96])
    _, right = msm.eval()
    assert right == oracle.msm(pts, sc, n)


def test_fixed_base_mul_montgomery_scalars(gpu, oracle):
    """the same scalars in Montgomery form (the in-memory form of midnight_curves::Fq) give the same points"""
    H = gpu.host
    vals = [3, R - 5, 0x1234567890ABCDEF << 100]
    sc = H.FrVec.from_ints(vals)                       # Montgomery form on the device
    out = H.DeviceBuffer(96 * len(vals))
    gpu.capi.check(gpu.lib().b200zk_g1_fixed_mul_dev(sc.ptr, gpu.FMT_MONT, len(vals), out.ptr, None))
    got = H.g1_export(out, len(vals))
    G = oracle.g1_generator()
    for i, v in enumerate(vals):
        assert got[96 * i:96 * i + 96] == oracle.g1_mul(G, fr(v)), hex(v)
