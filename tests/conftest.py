"""Shared fixtures.  `-m "not gpu"` tests run on a CPU-only box; `-m gpu` tests are the parity
tests proper and call the CUDA path through the C ABI."""
import ctypes as C
import importlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _build_oracle():
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "b200zk_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return so


class Oracle:
    """ctypes view of oracle/liboracle.so (the checker)."""

    def __init__(self, path):
        L = C.CDLL(path)
        L.orc_init()
        self.L = L
        u64 = C.c_uint64
        L.orc_g1_msm.argtypes = [C.c_char_p, C.c_char_p, u64, C.c_char_p, C.c_int]
        L.orc_g1_msm_naive.argtypes = [C.c_char_p, C.c_char_p, u64, C.c_char_p]
        L.orc_g1_sum.argtypes = [C.c_char_p, u64, C.c_char_p]
        L.orc_g1_synth_bases.argtypes = [u64, u64, u64, C.c_char_p, C.c_int]
        L.orc_fr_synth.argtypes = [u64, u64, u64, C.c_char_p]
        L.orc_fr_dot_u64.argtypes = [C.c_char_p, C.c_void_p, u64, C.c_char_p]
        L.orc_ntt.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_char_p, C.c_int]
        L.orc_ntt_naive.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.orc_fr_kate_div.argtypes = [C.c_char_p, u64, C.c_char_p, C.c_char_p, C.c_char_p]
        L.orc_fr_running_product.argtypes = [C.c_char_p, u64, C.c_char_p, C.c_int, C.c_char_p]
        L.orc_fr_batch_invert.argtypes = [C.c_char_p, u64, C.c_char_p]
        L.orc_fr_lincomb.argtypes = [C.c_char_p, C.c_char_p, u64, u64, C.c_char_p]

    def msm(self, points: bytes, scalars: bytes, n: int, naive=False) -> bytes:
        out = C.create_string_buffer(96)
        if naive:
            self.L.orc_g1_msm_naive(points, scalars, n, out)
        else:
            self.L.orc_g1_msm(points, scalars, n, out, 0)
        return out.raw

    def g1_sum(self, points: bytes, n: int) -> bytes:
        out = C.create_string_buffer(96)
        self.L.orc_g1_sum(points, n, out)
        return out.raw

    def g1_mul(self, point: bytes, scalar: bytes) -> bytes:
        out = C.create_string_buffer(96)
        self.L.orc_g1_mul(point, scalar, out)
        return out.raw

    def g1_generator(self) -> bytes:
        out = C.create_string_buffer(96)
        self.L.orc_g1_generator(out)
        return out.raw

    def g1_compress(self, p: bytes) -> bytes:
        out = C.create_string_buffer(48)
        self.L.orc_g1_compress(p, out)
        return out.raw

    def g1_decompress(self, c: bytes):
        out = C.create_string_buffer(96)
        rc = self.L.orc_g1_decompress(c, out)
        return rc, out.raw

    def synth_bases(self, seed, start, n) -> bytes:
        out = C.create_string_buffer(96 * n)
        self.L.orc_g1_synth_bases(seed, start, n, out, 0)
        return out.raw

    def synth_scalars(self, seed, start, n) -> bytes:
        out = C.create_string_buffer(32 * n)
        self.L.orc_fr_synth(seed, start, n, out)
        return out.raw

    def ntt(self, data: bytes, log_n, omega: bytes, flags=0, coset_in=None, coset_out=None) -> bytes:
        buf = C.create_string_buffer(data, len(data))
        self.L.orc_ntt(buf, log_n, omega, flags, coset_in, coset_out, 0)
        return buf.raw

    def ntt_naive(self, data: bytes, log_n, omega: bytes) -> bytes:
        buf = C.create_string_buffer(data, len(data))
        self.L.orc_ntt_naive(buf, log_n, omega)
        return buf.raw

    def kate_div(self, coeffs: bytes, z: bytes):
        n = len(coeffs) // 32
        quot = C.create_string_buffer(32 * max(n - 1, 1))
        ev = C.create_string_buffer(32)
        self.L.orc_fr_kate_div(coeffs, n, z, quot, ev)
        return quot.raw[:32 * max(n - 1, 0)], ev.raw

    def running_product(self, v: bytes, init: bytes = None, inclusive=False) -> bytes:
        out = C.create_string_buffer(max(len(v), 1))
        self.L.orc_fr_running_product(v, len(v) // 32, init, 1 if inclusive else 0, out)
        return out.raw[:len(v)]

    def batch_invert(self, v: bytes) -> bytes:
        out = C.create_string_buffer(max(len(v), 1))
        self.L.orc_fr_batch_invert(v, len(v) // 32, out)
        return out.raw[:len(v)]

    def lincomb(self, polys: bytes, coeffs: bytes, count: int) -> bytes:
        n = len(polys) // 32 // count
        out = C.create_string_buffer(32 * max(n, 1))
        self.L.orc_fr_lincomb(polys, coeffs, count, n, out)
        return out.raw[:32 * n]

    def field(self, name, a: bytes, b: bytes = None) -> bytes:
        n = 48 if name.startswith("orc_fp") else 32
        out = C.create_string_buffer(n)
        fn = getattr(self.L, name)
        if b is None:
            fn(a, out)
        else:
            fn(a, b, out)
        return out.raw


@pytest.fixture(scope="session")
def oracle():
    return Oracle(_build_oracle())


@pytest.fixture(scope="session")
def pyref():
    import pyref as P
    return P


@pytest.fixture(scope="session")
def kats():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))


@pytest.fixture(scope="session")
def zk():
    """The product package (directory name has a hyphen, hence import_module)."""
    return importlib.import_module("plutus-halo2-verifier-gen_b200")


@pytest.fixture(scope="session")
def gpu(zk):
    """Initialised CUDA backend; GPU tests fail (not skip) if the extension cannot run."""
    zk.init(-1)
    return zk
