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


def test_kzg_coefficient_and_lagrange_commitments_agree(gpu, oracle, pyref):
    """The identity the prover relies on when it mixes commit and commit_lagrange: with g_i = s^i G and
    g_lagrange_i = L_i(s) G (what ParamsKZG::unsafe_setup produces, /root/reference/src/kzg_params.rs:43),
    commit(coefficients) == commit_lagrange(NTT(coefficients)) == p(s) G.  Ties the CUDA MSM and the
    CUDA NTT together through a value neither of them computes."""
    k = 7
    n = 1 << k
    s = 0x1234567890ABCDEF1234567890ABCDEF % R
    G = oracle.g1_generator()
    omega = pyref.omega(k)
    powers = [pow(s, i, R) for i in range(n)]
    # L_i(s) = (s^n - 1)/n * w^i / (s - w^i)
    zn = (pow(s, n, R) - 1) * pyref.fr_inv(n) % R
    lag = [zn * pow(omega, i, R) % R * pyref.fr_inv((s - pow(omega, i, R)) % R) % R for i in range(n)]
    g = b"".join(oracle.g1_mul(G, fr(x)) for x in powers)
    gl = b"".join(oracle.g1_mul(G, fr(x)) for x in lag)
    params = gpu.host.ParamsKZG(k, g, gl)
    dom = gpu.host.EvaluationDomain(4, k)
    coeffs = oracle.synth_scalars(77, 0, n)
    evals = dom.coeff_to_lagrange(coeffs)
    c1 = gpu.host.KZGCommitmentScheme.commit(params, coeffs)
    c2 = gpu.host.KZGCommitmentScheme.commit_lagrange(params, evals)
    ps = sum(int.from_bytes(coeffs[32 * i:32 * i + 32], "little") * powers[i] for i in range(n)) % R
    assert c1 == c2 == oracle.g1_mul(G, fr(ps))
    # and the verifier-side shape: e(pi, [s]G2) == e(right, G2) holds iff s*left == right in G1 when s is known
    # (opening of p at z with quotient q = (p - p(z)) / (X - z)): commit(p) - p(z) G + z*commit(q) == s*commit(q)
    z = 0xDEADBEEF
    cz = [int.from_bytes(coeffs[32 * i:32 * i + 32], "little") for i in range(n)]
    pz = 0
    q = [0] * n
    for i in range(n - 1, -1, -1):        # synthetic division by (X - z)
        q[i] = pz
        pz = (pz * z + cz[i]) % R
    qb = b"".join(fr(x) for x in q)
    pi = gpu.host.KZGCommitmentScheme.commit(params, qb)
    dual = gpu.host.DualMSM()
    dual.append_left(1, pi)
    dual.append_right(1, c1)
    dual.append_right(-pz, G)
    dual.append_right(z, pi)
    left, right = dual.eval()
    assert oracle.g1_mul(left, fr(s)) == right
    params.release()


def test_streamed_column_batch(gpu):
    """A batched commitment with a long transfer (>= 128 MiB of scalars, e.g. 18 columns at k = 19) goes up in two groups of
    columns, the second copy running under the first group's MSMs.  A child process lowers the threshold so that a small
    batch takes that path: every column on both sides of the split must match the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import ctypes as C, importlib, os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
from conftest import Oracle, _build_oracle
zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
zk.init(-1)
orc = Oracle(_build_oracle())
for k, batch in ((10, 4), (12, 9), (13, 16)):
    n = 1 << k
    g = orc.synth_bases(0xB200, 0, n)
    params = zk.host.ParamsKZG(k, g)
    cols = [orc.synth_scalars(600 + i, 0, n) for i in range(batch)]
    got = zk.host.KZGCommitmentScheme.commit_batch(params, cols)
    for i in range(batch):
        assert got[i] == orc.msm(g, cols[i], n), (k, i)
    params.release()
print("streamed ok")
''' % (root, root)
    env = dict(os.environ, B200ZK_BATCH_STREAM_MIN_BYTES="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "streamed ok" in r.stdout, r.stdout + r.stderr
