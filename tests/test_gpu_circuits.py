"""Parity on the shapes of the reference's example circuits (BASELINE configs 0, 1, 3, 4): every
commitment of a proof-sized batch and every domain transform of the trace, bit-exact against the
oracle, with prover-like scalar columns."""
import ctypes as C
import random

import pytest

pytestmark = pytest.mark.gpu

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def fr(x):
    return (x % R).to_bytes(32, "little")


def prover_like_column(oracle, seed, n):
    rnd = random.Random(seed)
    uni = oracle.synth_scalars(seed, 0, n)
    vals = []
    for i in range(n):
        u = rnd.random()
        if i >= n - 6 or u >= 0.9:
            vals.append(uni[32 * i:32 * i + 32])      # blinding rows and the occasional full-width value
        elif u < 0.7:
            vals.append(fr(rnd.randrange(2)))
        else:
            vals.append(fr(rnd.randrange(1 << 16)))
    return b"".join(vals)


@pytest.mark.parametrize("name,k,ncom,nuni", [("simple_mul", 5, 10, 5), ("lookup_table", 11, 20, 6), ("atms", 14, 18, 6)])
def test_proof_shaped_commitment_batch(gpu, oracle, name, k, ncom, nuni):
    n = 1 << k
    g = oracle.synth_bases(0xB200, 0, n)
    gl = oracle.synth_bases(0xB201, 0, n)
    params = gpu.host.ParamsKZG(k, g, gl)
    advice = [prover_like_column(oracle, 300 + i, n) for i in range(ncom - nuni)]
    quotient = [oracle.synth_scalars(400 + i, 0, n) for i in range(nuni)]
    K = gpu.host.KZGCommitmentScheme
    got = K.commit_batch(params, advice, lagrange=True) + K.commit_batch(params, quotient)
    exp = [oracle.msm(gl, p, n) for p in advice] + [oracle.msm(g, p, n) for p in quotient]
    assert got == exp
    # compressed form is what the transcript absorbs: check the encoder on real commitments too
    for p in got[:3]:
        assert gpu.host.g1_compress(p) == oracle.g1_compress(p)
    params.release()


@pytest.mark.parametrize("k,j", [(5, 4), (11, 5), (14, 4)])
def test_domain_transform_trace(gpu, oracle, pyref, k, j):
    d = gpu.host.EvaluationDomain(j, k)
    n = d.n
    cols = [prover_like_column(oracle, 500 + i, n) for i in range(4)]
    coeffs = d.lagrange_to_coeff_batch(cols)
    for c_, l_ in zip(coeffs, cols):
        assert c_ == oracle.ntt(l_, k, fr(d.omega_inv), 1)
    ext = d.coeff_to_extended(coeffs[0])
    assert ext == oracle.ntt(coeffs[0] + bytes(32 * ((1 << d.extended_k) - n)), d.extended_k, fr(d.extended_omega), 0, fr(d.g_coset))
    assert d.extended_to_coeff(ext)[:32 * n] == coeffs[0]


def test_batched_verifier_full_shape(gpu, oracle):
    """config 5: the right-hand MSM of a 1024-proof batch (1024 x 26 points, not resident)"""
    n = 1024 * 26
    pts = oracle.synth_bases(0xB600, 0, n)
    sc = oracle.synth_scalars(31, 0, n)
    out = C.create_string_buffer(96)
    gpu.capi.check(gpu.lib().b200zk_msm_g1_adhoc(gpu.capi.addr(pts), 0, gpu.capi.addr(sc), 0, n, gpu.capi.addr(out)))
    assert out.raw == oracle.msm(pts, sc, n)
