"""Parity at the sizes BASELINE.json names: G1 MSM at 2^22 (three scalar distributions) and 2^24 in full against the C
oracle and through the discrete-log identity, and every commitment / domain transform of the proof-shaped traces at
k = 17 and k = 19 (atms, atms_with_lookups, sha256 shapes; configs[3] and configs[4])."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SEED = 0xB200


def fr(x):
    return (x % R).to_bytes(32, "little")


def splitmix64(x):
    x = x + np.uint64(0x9E3779B97F4A7C15)
    z = x.copy()
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def scalars(oracle, kind, seed, n):
    """U: uniform mod r (the checker's generator); S: prover-like (70 % bits, 20 % < 2^16, 10 % uniform, six blinding rows);
    A: adversarial, every scalar r - 1."""
    uni = np.frombuffer(oracle.synth_scalars(seed, 0, n), dtype=np.uint64).reshape(n, 4).copy()
    if kind == "U":
        return uni
    if kind == "A":
        uni[:] = np.frombuffer(fr(R - 1), dtype=np.uint64)
        return uni
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    m1, m2 = u < 0.7, (u >= 0.7) & (u < 0.9)
    m1[-6:] = False
    m2[-6:] = False
    uni[m1, 0] = rng.integers(0, 2, n, dtype=np.uint64)[m1]
    uni[m2, 0] = rng.integers(0, 1 << 16, n, dtype=np.uint64)[m2]
    uni[m1 | m2, 1:] = 0
    return uni


def device_table(gpu, n, seed=SEED, flags=0):
    """Synthetic SRS table generated on the GPU, registered, and read back for the checker."""
    import torch
    d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
    gpu.capi.check(gpu.lib().b200zk_g1_synth_bases_dev(seed, 0, n, d_b.data_ptr(), None))
    torch.cuda.synchronize()
    h = C.c_uint64(0)
    gpu.capi.check(gpu.lib().b200zk_bases_register_dev(d_b.data_ptr(), n, gpu.FMT_MONT | flags, 96, C.byref(h)))
    del d_b
    pts = np.empty(96 * n, dtype=np.uint8)
    gpu.capi.check(gpu.lib().b200zk_bases_read(h.value, 0, n, pts.ctypes.data))
    return h.value, pts


def oracle_msm(oracle, pts, sc, n):
    out = C.create_string_buffer(96)
    oracle.L.orc_g1_msm(pts.ctypes.data_as(C.c_char_p), sc.ctypes.data_as(C.c_char_p), n, out, 0)
    return out.raw


def dlog_point(oracle, sc, seed, n):
    a = np.ascontiguousarray(splitmix64(np.arange(n, dtype=np.uint64) + np.uint64(seed)))
    dot = C.create_string_buffer(32)
    oracle.L.orc_fr_dot_u64(sc.ctypes.data_as(C.c_char_p), a.ctypes.data, n, dot)
    return oracle.g1_mul(oracle.g1_generator(), dot.raw)


def gpu_msm(gpu, h, sc, n, batch=1):
    out = C.create_string_buffer(96 * batch)
    gpu.capi.check(gpu.lib().b200zk_msm_g1_batch(h, 0, sc.ctypes.data, n, batch, 0, gpu.capi.addr(out)))
    return [out.raw[96 * i:96 * i + 96] for i in range(batch)]


@pytest.fixture(scope="module")
def table22(gpu, oracle):
    n = 1 << 22
    h, pts = device_table(gpu, n)
    # the bases themselves: a sample against the checker's generator
    for start in (0, 1 << 21, n - 8):
        assert pts[96 * start:96 * (start + 8)].tobytes() == oracle.synth_bases(SEED, start, 8)
    yield h, pts, n
    gpu.capi.check(gpu.lib().b200zk_bases_release(h))


@pytest.mark.parametrize("kind", ["U", "S", "A"])
def test_msm_2pow22_vs_oracle(gpu, oracle, table22, kind):
    h, pts, n = table22
    sc = scalars(oracle, kind, 71, n)
    got = gpu_msm(gpu, h, sc, n)[0]
    assert got == oracle_msm(oracle, pts, sc, n)
    assert got == dlog_point(oracle, sc, SEED, n)


def test_msm_2pow24_headline_config(gpu, oracle):
    """BASELINE's headline size: the uniform column in full against the C oracle (over the bases read back from the GPU's
    table) and by the discrete-log identity; the prover-like and the adversarial column by the identity; resident and
    host-buffer (streamed) entry points agree."""
    import torch
    n = 1 << 24
    h, pts = device_table(gpu, n)
    sc = scalars(oracle, "U", 1, n)
    got = gpu_msm(gpu, h, sc, n)[0]
    assert got == dlog_point(oracle, sc, SEED, n)
    assert got == oracle_msm(oracle, pts, sc, n)
    del pts
    d_sc = torch.from_numpy(sc.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    gpu.capi.check(gpu.lib().b200zk_msm_g1_dev(h, 0, d_sc.data_ptr(), n, 1, 0, None, d_out.data_ptr(), None))
    torch.cuda.synchronize()
    assert bytes(d_out.cpu().numpy()) == got
    for kind in ("S", "A"):
        sc = scalars(oracle, kind, 9, n)
        assert gpu_msm(gpu, h, sc, n)[0] == dlog_point(oracle, sc, SEED, n)
    gpu.capi.check(gpu.lib().b200zk_bases_release(h))


#                         name            k  commitments  of which uniform (quotient pieces, f, pi, random poly)
LARGE = [("atms", 17, 18, 6), ("atms_with_lookups", 17, 22, 6), ("atms", 19, 18, 6), ("sha256", 19, 26, 7)]


@pytest.mark.parametrize("name,k,ncom,nuni", LARGE)
def test_proof_shaped_commitment_batch_large(gpu, oracle, name, k, ncom, nuni):
    """Every commitment of a proof-sized batch at BASELINE configs[3] / configs[4] sizes, against the oracle; columns are passed
    one buffer each (pointer-array entry point), advice-like columns against the Lagrange table, the rest against g."""
    n = 1 << k
    hg, g = device_table(gpu, n, SEED)
    hl, gl = device_table(gpu, n, SEED + 1)
    advice = [scalars(oracle, "S", 300 + i, n) for i in range(ncom - nuni)]
    uniform = [scalars(oracle, "U", 400 + i, n) for i in range(nuni)]
    for h, pts, cols in ((hl, gl, advice), (hg, g, uniform)):
        ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        out = C.create_string_buffer(96 * len(cols))
        gpu.capi.check(gpu.lib().b200zk_msm_g1_batch_ptrs(h, 0, C.addressof(ptrs), n, len(cols), 0, gpu.capi.addr(out)))
        for j, c in enumerate(cols):
            assert out.raw[96 * j:96 * j + 96] == oracle_msm(oracle, pts, c, n), (name, k, j)
    for h in (hg, hl):
        gpu.capi.check(gpu.lib().b200zk_bases_release(h))


@pytest.mark.parametrize("k,j", [(17, 4), (19, 5)])
def test_domain_transform_trace_large(gpu, oracle, k, j):
    d = gpu.host.EvaluationDomain(j, k)
    n = d.n
    cols = [scalars(oracle, "S", 500 + i, n).tobytes() for i in range(3)]
    coeffs = d.lagrange_to_coeff_batch(cols)
    for c_, l_ in zip(coeffs, cols):
        assert c_ == oracle.ntt(l_, k, fr(d.omega_inv), 1)
    ext = d.coeff_to_extended(coeffs[0])
    assert ext == oracle.ntt(coeffs[0] + bytes(32 * ((1 << d.extended_k) - n)), d.extended_k, fr(d.extended_omega), 0, fr(d.g_coset))
    assert d.extended_to_coeff(ext)[:32 * n] == coeffs[0]
    assert d.coeff_to_lagrange(coeffs[1]) == cols[1]
