"""Byte-exact parity of the CUDA Fr NTT (through the C ABI) with the oracle: every size 2^0..2^14,
the two-pass sizes, inverse / coset variants, batches, Montgomery-form data, and the BASELINE size
2^22; plus size-independent properties (round trip, linearity) at the large sizes."""
import ctypes as C
import random

import pytest

pytestmark = pytest.mark.gpu


def fr(x):
    return (x % 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001).to_bytes(32, "little")


def gpu_ntt(gpu, data: bytes, log_n, omega, flags=0, shift=None, batch=1) -> bytes:
    buf = bytearray(data)
    wb = fr(omega)                                   # keep the byte strings alive across the call
    sb = fr(shift) if shift is not None else None
    gpu.capi.check(gpu.lib().b200zk_ntt_fr_batch(gpu.capi.addr(buf), batch, log_n, gpu.capi.addr(wb), flags,
                                                 gpu.capi.addr(sb)))
    return bytes(buf)


@pytest.mark.parametrize("log_n", list(range(0, 15)) + [16, 17, 20])
def test_forward_inverse_coset_vs_oracle(gpu, oracle, pyref, log_n):
    data = oracle.synth_scalars(2 + log_n, 0, 1 << log_n)
    w = pyref.omega(log_n)
    wi = pyref.fr_inv(w)
    fwd = gpu_ntt(gpu, data, log_n, w)
    assert fwd == oracle.ntt(data, log_n, fr(w)), "forward"
    assert gpu_ntt(gpu, fwd, log_n, wi, gpu.NTT_INVERSE_SCALE) == data, "inverse"
    g = 7
    cos = gpu_ntt(gpu, data, log_n, w, gpu.NTT_COSET_IN, g)
    assert cos == oracle.ntt(data, log_n, fr(w), 0, fr(g)), "coset forward"
    back = gpu_ntt(gpu, cos, log_n, wi, gpu.NTT_INVERSE_SCALE | gpu.NTT_COSET_OUT, pyref.fr_inv(g))
    assert back == data, "coset inverse"
    assert back == oracle.ntt(cos, log_n, fr(wi), 1, None, fr(pyref.fr_inv(g)))


def test_small_sizes_vs_quadratic_dft(gpu, oracle, pyref):
    for log_n in range(0, 8):
        data = oracle.synth_scalars(99, 7, 1 << log_n)
        w = pyref.omega(log_n)
        assert gpu_ntt(gpu, data, log_n, w) == oracle.ntt_naive(data, log_n, fr(w))


@pytest.mark.parametrize("log_n,batch", [(3, 5), (5, 20), (8, 7), (11, 3), (12, 4), (14, 2)])
def test_batch(gpu, oracle, pyref, log_n, batch):
    n = 1 << log_n
    data = oracle.synth_scalars(5, 0, n * batch)
    w = pyref.omega(log_n)
    got = gpu_ntt(gpu, data, log_n, w, batch=batch)
    for b in range(batch):
        assert got[32 * n * b:32 * n * (b + 1)] == oracle.ntt(data[32 * n * b:32 * n * (b + 1)], log_n, fr(w)), b


def test_over_range_input_is_reduced(gpu, oracle, pyref):
    """canonical inputs >= r are reduced on read like the proof wire format (transcript.ak:158-179)"""
    log_n = 6
    vals = [pyref.R_MOD, pyref.R_MOD + 5, (1 << 256) - 1] + list(range(61))
    data = b"".join(v.to_bytes(32, "little") for v in vals)
    red = b"".join(fr(v) for v in vals)
    w = pyref.omega(log_n)
    assert gpu_ntt(gpu, data, log_n, w) == oracle.ntt(red, log_n, fr(w))


def test_montgomery_form_data(gpu, oracle, pyref):
    log_n = 10
    data = oracle.synth_scalars(8, 0, 1 << log_n)
    mont = b"".join(fr(int.from_bytes(data[32 * i:32 * i + 32], "little") << 256) for i in range(1 << log_n))
    w = pyref.omega(log_n)
    got = gpu_ntt(gpu, mont, log_n, w, gpu.NTT_MONT)
    exp = oracle.ntt(data, log_n, fr(w))
    exp_mont = b"".join(fr(int.from_bytes(exp[32 * i:32 * i + 32], "little") << 256) for i in range(1 << log_n))
    assert got == exp_mont


def test_2pow22_full_parity(gpu, oracle, pyref):
    """BASELINE size: 2^22 forward transform, byte-identical to the oracle, then round trip."""
    log_n = 22
    data = oracle.synth_scalars(2, 0, 1 << log_n)
    w = pyref.omega(log_n)
    fwd = gpu_ntt(gpu, data, log_n, w)
    assert fwd == oracle.ntt(data, log_n, fr(w))
    assert gpu_ntt(gpu, fwd, log_n, pyref.fr_inv(w), gpu.NTT_INVERSE_SCALE) == data


@pytest.mark.parametrize("log_n", [23, 24])
def test_2pow23_2pow24_full_parity(gpu, oracle, pyref, log_n):
    """The three-pass sizes: forward, coset and inverse, byte-identical to the oracle in full."""
    n = 1 << log_n
    data = oracle.synth_scalars(40 + log_n, 0, n)
    w = pyref.omega(log_n)
    fwd = gpu_ntt(gpu, data, log_n, w)
    assert fwd == oracle.ntt(data, log_n, fr(w))
    assert gpu_ntt(gpu, fwd, log_n, pyref.fr_inv(w), gpu.NTT_INVERSE_SCALE) == data
    if log_n == 23:
        cos = gpu_ntt(gpu, data, log_n, w, gpu.NTT_COSET_IN, 7)
        assert cos == oracle.ntt(data, log_n, fr(w), 0, fr(7))
        assert gpu_ntt(gpu, cos, log_n, pyref.fr_inv(w), gpu.NTT_INVERSE_SCALE | gpu.NTT_COSET_OUT, pyref.fr_inv(7)) == data
        # a batch of two, Montgomery data
        two = data[:32 * (n // 2)] * 2   # two polynomials of 2^22 each
        assert gpu_ntt(gpu, two, log_n - 1, pyref.omega(log_n - 1), 0, None, 2) == oracle.ntt(two[:32 * (n // 2)], log_n - 1, fr(pyref.omega(log_n - 1))) * 2


def test_2pow24_properties(gpu, oracle, pyref):
    """Largest bench size: inverse(forward(x)) == x and two outputs checked against sums of the input."""
    log_n = 24
    n = 1 << log_n
    data = oracle.synth_scalars(3, 0, n)
    w = pyref.omega(log_n)
    fwd = gpu_ntt(gpu, data, log_n, w)
    # X[0] = sum x_i ; X[n/2] = sum (-1)^i x_i
    import numpy as np
    s0 = s1 = 0
    arr = np.frombuffer(data, dtype="<u8").reshape(n, 4)
    for limb in range(4):
        col = arr[:, limb]
        lo, hi = int((col & 0xFFFFFFFF).sum(dtype=np.uint64)), int((col >> 32).sum(dtype=np.uint64))
        tot = lo + (hi << 32)
        ev = col[0::2]
        tot_even = int((ev & 0xFFFFFFFF).sum(dtype=np.uint64)) + (int((ev >> 32).sum(dtype=np.uint64)) << 32)
        s0 += tot << (64 * limb)
        s1 += (2 * tot_even - tot) << (64 * limb)
    assert fwd[:32] == fr(s0)
    assert fwd[32 * (n // 2):32 * (n // 2) + 32] == fr(s1)
    assert gpu_ntt(gpu, fwd, log_n, pyref.fr_inv(w), gpu.NTT_INVERSE_SCALE) == data


def test_evaluation_domain_mirror(gpu, oracle, pyref):
    """EvaluationDomain round trips with the reference's shapes (k = 10, j = 4 -> extended_k = 12)."""
    d = gpu.host.EvaluationDomain(4, 10, g_coset=gpu.host.ZETA)
    n = d.n
    lag = oracle.synth_scalars(11, 0, n)
    coeff = d.lagrange_to_coeff(lag)
    assert coeff == oracle.ntt(lag, d.k, fr(d.omega_inv), 1)
    assert d.coeff_to_lagrange(coeff) == lag
    ext = d.coeff_to_extended(coeff)
    padded = coeff + bytes(32 * ((1 << d.extended_k) - n))
    assert ext == oracle.ntt(padded, d.extended_k, fr(d.extended_omega), 0, fr(d.g_coset))
    back = d.extended_to_coeff(ext)
    assert len(back) == 32 * n * 3 and back[:32 * n] == coeff and back[32 * n:] == bytes(32 * n * 2)
    polys = [oracle.synth_scalars(12 + i, 0, n) for i in range(5)]
    assert d.lagrange_to_coeff_batch(polys) == [d.lagrange_to_coeff(p) for p in polys]


def test_argument_errors(gpu):
    data = bytearray(64)
    one = fr(1)
    L = gpu.lib()
    assert L.b200zk_ntt_fr(gpu.capi.addr(data), 40, gpu.capi.addr(one), 0, 0) == -1
    assert L.b200zk_ntt_fr(gpu.capi.addr(data), 1, gpu.capi.addr(one), gpu.NTT_COSET_IN, 0) == -1
    assert L.b200zk_ntt_fr(0, 1, gpu.capi.addr(one), 0, 0) == -1


def test_pipelined_host_batch(gpu, oracle, pyref):
    """Host-buffer batches of >= 32 MiB are pipelined group by group (upload of the next group and download of the previous
    one under the current group's kernels): 8 x 2^17 elements, every polynomial against the oracle, plus coset + inverse."""
    log_n, batch = 17, 8
    n = 1 << log_n
    data = oracle.synth_scalars(77, 0, n * batch)
    w = pyref.omega(log_n)
    got = gpu_ntt(gpu, data, log_n, w, batch=batch)
    for b in range(batch):
        assert got[32 * n * b:32 * n * (b + 1)] == oracle.ntt(data[32 * n * b:32 * n * (b + 1)], log_n, fr(w)), b
    cos = gpu_ntt(gpu, data, log_n, w, gpu.NTT_COSET_IN, 7, batch=batch)
    back = gpu_ntt(gpu, cos, log_n, pyref.fr_inv(w), gpu.NTT_INVERSE_SCALE | gpu.NTT_COSET_OUT, pyref.fr_inv(7), batch=batch)
    assert back == data
    assert cos[:32 * n] == oracle.ntt(data[:32 * n], log_n, fr(w), 0, fr(7))
