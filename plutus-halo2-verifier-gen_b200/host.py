"""Host-side mirror of the reference-facing interface for the MSM / NTT path.

The reference reaches this path only through upstream traits (the crate sources are not in
/root/reference, so the trait shapes below are restated from the reference's call sites):

* ``ParamsKZG``                       /root/reference/src/kzg_params.rs:33-80 (SRS cache; the MSM bases)
* ``KZGCommitmentScheme::commit`` /
  ``commit_lagrange``                 reached via create_proof / keygen_vk,
                                      /root/reference/examples/simple_mul.rs:62,72
* ``EvaluationDomain``                /root/reference/examples/ivc.rs:109,
                                      /root/reference/src/circuits/ivc_circuit.rs:305
* ``DualMSM`` (``Guard::verify``)     /root/reference/examples/simple_mul.rs:98-102, examples/ivc.rs:196;
                                      the two sums are spelled out at
                                      /root/reference/aiken-verifier/aiken_halo2/lib/halo2_kzg.ak:15-44

Everything numeric is done by libb200zk.so on the GPU; this file only marshals buffers and
computes domain constants (a handful of modular exponentiations with Python ints).
Polynomials are byte strings of 32-byte little-endian canonical Fr elements; points are
96-byte affine wire-format strings (see include/b200zk.h).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

from . import capi
from .capi import FMT_CANONICAL, NTT_COSET_IN, NTT_COSET_OUT, NTT_INVERSE_SCALE, addr, check, lib

# scalar field of BLS12-381 (plinth-verifier/plutus-halo2/src/Plutus/Crypto/BlsTypes.hs:97)
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
TWO_ADICITY = 32
MULTIPLICATIVE_GENERATOR = 7  # Constants.hs:10-13 (DELTA = 7^(2^32))
ROOT_OF_UNITY = pow(MULTIPLICATIVE_GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)
# a primitive cube root of unity; upstream's coset generator (F::ZETA) is one of the two --
# which one is not verifiable here, so EvaluationDomain takes g_coset as a parameter
ZETA = pow(MULTIPLICATIVE_GENERATOR, (R_MOD - 1) // 3, R_MOD)


def fr_bytes(x: int) -> bytes:
    return (x % R_MOD).to_bytes(32, "little")


class ParamsKZG:
    """Device-resident SRS: ``g`` (monomial basis) and ``g_lagrange`` (Lagrange basis) tables.

    Mirrors the two base tables of upstream ``ParamsKZG<Bls12>`` that
    ``get_or_create_kzg_params`` returns (/root/reference/src/kzg_params.rs:33-47)."""

    def __init__(self, k: int, g: bytes, g_lagrange: Optional[bytes] = None, fmt: int = FMT_CANONICAL,
                 stride: int = 96, flags: int = 0):
        """``flags``: capi.BASES_* bits OR-ed into the format (window tables off, shard / replicate over the bound GPUs)."""
        self.k = k
        self.n = 1 << k
        fmt |= flags
        self._g = self._register(g, fmt, stride)
        self._g_lagrange = self._register(g_lagrange, fmt, stride) if g_lagrange is not None else None

    def _register(self, table: bytes, fmt: int, stride: int) -> int:
        if len(table) < (self.n - 1) * stride + 96:
            raise ValueError("SRS table shorter than 2^k points")
        h = C.c_uint64(0)
        check(lib().b200zk_bases_register(addr(table), self.n, fmt, stride, C.byref(h)))
        return h.value

    @classmethod
    def from_device(cls, k: int, d_g: int, d_g_lagrange: Optional[int] = None, fmt: int = capi.FMT_MONT):
        """Adopt tables that already live in HBM (e.g. produced by the synthetic generator)."""
        self = cls.__new__(cls)
        self.k, self.n = k, 1 << k
        h = C.c_uint64(0)
        check(lib().b200zk_bases_register_dev(d_g, self.n, fmt, 96, C.byref(h)))
        self._g = h.value
        self._g_lagrange = None
        if d_g_lagrange is not None:
            check(lib().b200zk_bases_register_dev(d_g_lagrange, self.n, fmt, 96, C.byref(h)))
            self._g_lagrange = h.value
        return self

    def release(self) -> None:
        for h in (self._g, self._g_lagrange):
            if h:
                check(lib().b200zk_bases_release(h))
        self._g = self._g_lagrange = None


class KZGCommitmentScheme:
    """``commit`` / ``commit_lagrange`` of upstream ``KZGCommitmentScheme<Bls12>``: one G1 MSM of
    the polynomial's n coefficients (or evaluations) against the matching SRS table.  Like
    upstream, a polynomial longer than the table is a programming error (asserted)."""

    @staticmethod
    def _msm(handle: int, n_max: int, poly: bytes, batch: int = 1) -> List[bytes]:
        if len(poly) % (32 * batch):
            raise ValueError("polynomial bytes must be batch * n * 32")
        n = len(poly) // (32 * batch)
        assert n <= n_max, "polynomial longer than the SRS"
        out = C.create_string_buffer(96 * batch)
        check(lib().b200zk_msm_g1_batch(handle, 0, addr(poly), n, batch, FMT_CANONICAL, addr(out)))
        return [out.raw[96 * i:96 * (i + 1)] for i in range(batch)]

    @staticmethod
    def commit(params: ParamsKZG, poly: bytes) -> bytes:
        return KZGCommitmentScheme._msm(params._g, params.n, poly)[0]

    @staticmethod
    def commit_lagrange(params: ParamsKZG, poly: bytes) -> bytes:
        if params._g_lagrange is None:
            raise ValueError("params were built without a Lagrange-basis table")
        return KZGCommitmentScheme._msm(params._g_lagrange, params.n, poly)[0]

    @staticmethod
    def commit_batch(params: ParamsKZG, polys: Sequence[bytes], lagrange: bool = False) -> List[bytes]:
        """All columns of one prover phase in a single call (same length each).  The columns stay where they are (one
        buffer per polynomial, like halo2's Vecs): the library takes an array of pointers and, with several GPUs bound,
        deals the columns out."""
        if not polys:
            return []
        h = params._g_lagrange if lagrange else params._g
        if h is None:
            raise ValueError("params were built without a Lagrange-basis table")
        n = len(polys[0]) // 32
        if any(len(p) != 32 * n for p in polys):
            raise ValueError("columns of one batch have the same length")
        assert n <= params.n, "polynomial longer than the SRS"
        ptrs = (C.c_void_p * len(polys))(*[addr(p) for p in polys])
        out = C.create_string_buffer(96 * len(polys))
        check(lib().b200zk_msm_g1_batch_ptrs(h, 0, C.addressof(ptrs), n, len(polys), FMT_CANONICAL, addr(out)))
        return [out.raw[96 * i:96 * (i + 1)] for i in range(len(polys))]


class EvaluationDomain:
    """Mirror of upstream ``EvaluationDomain::new(j, k)``: n = 2^k rows, quotient degree j - 1,
    extended domain of size 2^extended_k with extended_k = k + ceil(log2(j - 1)); omega follows
    the convention pinned at /root/reference/aiken-verifier/aiken_halo2/lib/omega_rotations.ak:48-81."""

    def __init__(self, j: int, k: int, g_coset: int = ZETA):
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ext = k
        while (1 << ext) < self.n * self.quotient_poly_degree:
            ext += 1
        if ext > TWO_ADICITY:
            raise ValueError("extended_k exceeds the 2-adicity of the field")
        self.extended_k = ext
        self.omega = pow(ROOT_OF_UNITY, 1 << (TWO_ADICITY - k), R_MOD)
        self.omega_inv = pow(self.omega, R_MOD - 2, R_MOD)
        self.extended_omega = pow(ROOT_OF_UNITY, 1 << (TWO_ADICITY - ext), R_MOD)
        self.extended_omega_inv = pow(self.extended_omega, R_MOD - 2, R_MOD)
        self.g_coset = g_coset % R_MOD
        self.g_coset_inv = pow(self.g_coset, R_MOD - 2, R_MOD)

    def get_omega(self) -> int:
        return self.omega

    def get_omega_inv(self) -> int:
        return self.omega_inv

    def get_extended_omega(self) -> int:
        return self.extended_omega

    @staticmethod
    def _ntt(data: bytearray, log_n: int, omega: int, flags: int, shift: Optional[int] = None, batch: int = 1):
        sh = fr_bytes(shift) if shift is not None else None
        om = fr_bytes(omega)  # named so the buffer outlives the call
        check(lib().b200zk_ntt_fr_batch(addr(data), batch, log_n, addr(om), flags, addr(sh)))

    def lagrange_to_coeff(self, a: bytes) -> bytes:
        """Evaluations over H (n values) -> coefficients (inverse NTT, scaled by 1/n)."""
        assert len(a) == 32 * self.n
        buf = bytearray(a)
        self._ntt(buf, self.k, self.omega_inv, NTT_INVERSE_SCALE)
        return bytes(buf)

    def coeff_to_lagrange(self, a: bytes) -> bytes:
        assert len(a) == 32 * self.n
        buf = bytearray(a)
        self._ntt(buf, self.k, self.omega, 0)
        return bytes(buf)

    def coeff_to_extended(self, a: bytes) -> bytes:
        """n coefficients -> evaluations over the coset g_coset * H_ext (zero-padded, scaled by
        g_coset^i, forward NTT of size 2^extended_k)."""
        assert len(a) == 32 * self.n
        buf = bytearray(a) + bytearray(32 * ((1 << self.extended_k) - self.n))
        self._ntt(buf, self.extended_k, self.extended_omega, NTT_COSET_IN, self.g_coset)
        return bytes(buf)

    def extended_to_coeff(self, a: bytes) -> bytes:
        """Inverse of coeff_to_extended; returns the n * (j - 1) low coefficients like upstream."""
        assert len(a) == 32 << self.extended_k
        buf = bytearray(a)
        self._ntt(buf, self.extended_k, self.extended_omega_inv, NTT_INVERSE_SCALE | NTT_COSET_OUT, self.g_coset_inv)
        return bytes(buf[: 32 * self.n * self.quotient_poly_degree])

    def lagrange_to_coeff_batch(self, polys: Sequence[bytes]) -> List[bytes]:
        """Many columns at once, each in its own buffer (pointer-array entry point; dealt out over the bound GPUs)."""
        bufs = [bytearray(p) for p in polys]
        if not bufs:
            return []
        ptrs = (C.c_void_p * len(bufs))(*[addr(b) for b in bufs])
        om = fr_bytes(self.omega_inv)
        check(lib().b200zk_ntt_fr_batch_ptrs(C.addressof(ptrs), len(bufs), self.k, addr(om), NTT_INVERSE_SCALE, None))
        return [bytes(b) for b in bufs]


class DualMSM:
    """The verifier's pair of lazy MSMs (upstream ``DualMSM``; restated in-tree at
    /root/reference/aiken-verifier/aiken_halo2/lib/halo2_kzg.ak:15-44 and
    /root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/Halo2MultiOpenMSM.hs:60-98).

    ``eval`` evaluates both sums on the GPU.  The accept decision is
    ``e(left, [s]G2) == e(right, G2)`` (/root/reference/aiken-verifier/templates/verification_h2.hbs:121-128);
    the pairing itself stays with the caller's pairing library and is out of this path's scope."""

    def __init__(self):
        self.left: List[Tuple[int, bytes]] = []
        self.right: List[Tuple[int, bytes]] = []

    def append_left(self, scalar: int, point: bytes) -> None:
        self.left.append((scalar % R_MOD, point))

    def append_right(self, scalar: int, point: bytes) -> None:
        self.right.append((scalar % R_MOD, point))

    def scale(self, factor: int) -> None:
        self.left = [(s * factor % R_MOD, p) for s, p in self.left]
        self.right = [(s * factor % R_MOD, p) for s, p in self.right]

    def add_msm(self, other: "DualMSM") -> None:
        """batch_verify's random linear combination appends the other guard's terms
        (/root/reference/src/circuits/schnorr_circuit.rs:224)."""
        self.left += other.left
        self.right += other.right

    @staticmethod
    def _eval(terms: List[Tuple[int, bytes]]) -> bytes:
        out = C.create_string_buffer(96)
        if not terms:
            return out.raw
        pts = b"".join(p for _, p in terms)
        sc = b"".join(fr_bytes(s) for s, _ in terms)
        check(lib().b200zk_msm_g1_adhoc(addr(pts), FMT_CANONICAL, addr(sc), FMT_CANONICAL, len(terms), addr(out)))
        return out.raw

    def eval(self) -> Tuple[bytes, bytes]:
        return self._eval(self.left), self._eval(self.right)


class Transcript:
    """The reference's Fiat-Shamir transcript (CardanoFriendlyBlake2b, /root/reference/src/plutus_gen/adjusted_types/mod.rs:30-72;
    verifier twin /root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:19-106), kept inside the library so that
    ``multi_open`` / ``multi_prepare`` advance it in the same call that does the work.  Scalars are ints, points 48-byte
    compressed strings.  With ``proof`` given it also reads: ``read_point`` / ``read_scalar`` consume the next item and absorb it."""

    def __init__(self, proof: bytes = b""):
        h = C.c_uint64(0)
        check(lib().b200zk_transcript_new(C.byref(h)))
        self.handle = h.value
        self.proof = bytes(proof)
        self.pos = 0

    def common_scalar(self, s: int) -> None:
        b = fr_bytes(s)
        check(lib().b200zk_transcript_common_scalar(self.handle, addr(b)))

    def common_point(self, compressed: bytes) -> None:
        assert len(compressed) == 48
        check(lib().b200zk_transcript_common_point(self.handle, addr(compressed)))

    write_scalar, write_point = common_scalar, common_point

    def read_scalar(self) -> int:
        b = self.proof[self.pos:self.pos + 32]
        self.pos += 32
        check(lib().b200zk_transcript_common_scalar(self.handle, addr(b)))
        return int.from_bytes(b, "little")

    def read_point(self) -> bytes:
        b = self.proof[self.pos:self.pos + 48]
        self.pos += 48
        self.common_point(b)
        return b

    def squeeze_challenge(self) -> int:
        out = C.create_string_buffer(32)
        check(lib().b200zk_transcript_squeeze(self.handle, addr(out)))
        return int.from_bytes(out.raw, "little")

    def free(self) -> None:
        if self.handle:
            check(lib().b200zk_transcript_free(self.handle))
            self.handle = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().b200zk_transcript_free(self.handle)
        except Exception:
            pass


class Guard:
    """upstream ``DualMSM`` guard of one opening proof: accept iff e(left, [s]G2) == e(right, G2)."""

    def __init__(self, handle: int, scalars: Optional[dict] = None):
        self.handle = handle
        self.scalars = scalars or {}

    def eval(self) -> Tuple[bytes, bytes]:
        return batch_guards([self])

    def free(self) -> None:
        if self.handle:
            check(lib().b200zk_guard_free(self.handle))
            self.handle = None


def batch_guards(guards: Sequence[Guard], challenges: Optional[Sequence[int]] = None) -> Tuple[bytes, bytes]:
    """(left, right) of sum_i c_i * guard_i: ``batch_verify`` up to the pairing
    (/root/reference/src/circuits/schnorr_circuit.rs:224-229).  One GPU decompression batch + two ad-hoc MSMs."""
    hs = (C.c_uint64 * len(guards))(*[g.handle for g in guards])
    cb = b"".join(fr_bytes(c) for c in challenges) if challenges is not None else None
    left, right = C.create_string_buffer(96), C.create_string_buffer(96)
    check(lib().b200zk_guard_eval(C.addressof(hs), len(guards), addr(cb), addr(left), addr(right)))
    return left.raw, right.raw


def multi_open(params: "ParamsKZG", transcript: Transcript, polys: Sequence["FrVec"], queries: Sequence[Tuple[int, int]]) -> bytes:
    """``KZGCommitmentScheme::multi_open`` with resident polynomials: ``queries`` = (polynomial index, point); returns the opening
    proof f || q_evals || pi and leaves the transcript where the verifier's will be."""
    n = polys[0].n
    ptrs = (C.c_void_p * len(polys))(*[p.ptr for p in polys])
    qp = (C.c_uint32 * len(queries))(*[q[0] for q in queries])
    pts = b"".join(fr_bytes(q[1]) for q in queries)
    out = C.create_string_buffer(48 + 32 * len(queries) + 48)
    ln = C.c_size_t(0)
    check(lib().b200zk_h2mo_open_dev(params._g, transcript.handle, C.addressof(ptrs), len(polys), n, C.addressof(qp), addr(pts),
                                     len(queries), addr(out), len(out), C.byref(ln)))
    return out.raw[:ln.value]


def multi_prepare(transcript: Transcript, commitments: Sequence[bytes], queries: Sequence[Tuple[int, int, int]], proof: bytes) -> Guard:
    """``KZGCommitmentScheme::multi_prepare``: ``commitments`` 48-byte compressed points, ``queries`` = (commitment index, point,
    claimed evaluation), ``proof`` the opening proof.  Returns the guard; ``guard.scalars`` holds x1..x4, f_eval, v."""
    cb = b"".join(commitments)
    qc = (C.c_uint32 * len(queries))(*[q[0] for q in queries])
    pts = b"".join(fr_bytes(q[1]) for q in queries)
    evs = b"".join(fr_bytes(q[2]) for q in queries)
    g = C.c_uint64(0)
    sc = C.create_string_buffer(192)
    check(lib().b200zk_h2mo_prepare(transcript.handle, addr(cb), len(commitments), C.addressof(qc), addr(pts), addr(evs), len(queries),
                                    addr(proof), len(proof), C.byref(g), addr(sc)))
    names = ("x1", "x2", "x3", "x4", "f_eval", "v")
    return Guard(g.value, {nm: int.from_bytes(sc.raw[32 * i:32 * i + 32], "little") for i, nm in enumerate(names)})


def g1_compress(affine: bytes) -> bytes:
    """48-byte ZCash compressed form, what the transcript absorbs
    (/root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:62-83)."""
    out = C.create_string_buffer(48)
    check(lib().b200zk_g1_compress(addr(affine), addr(out)))
    return out.raw


def g1_decompress_batch(compressed: bytes, strict: bool = True) -> Tuple[bytes, List[int]]:
    """Decompresses n 48-byte points on the GPU (the proof's commitments, transcript.ak:62-83).
    Returns (n * 96 bytes of affine wire format, per-point status).  With ``strict`` a bad encoding
    raises like the in-tree uncompress does (CompressUncompress.hs:70-100); otherwise the caller reads
    the status list (0 ok, 1 not compressed, 2 bad infinity, 3 x >= p, 4 not on the curve)."""
    if len(compressed) % 48:
        raise ValueError("compressed points are 48 bytes each")
    n = len(compressed) // 48
    out = C.create_string_buffer(96 * max(n, 1))
    status = (C.c_uint32 * max(n, 1))()
    rc = lib().b200zk_g1_decompress_batch(addr(compressed), n, addr(out), C.addressof(status))
    if rc != 0 and (strict or rc != -7):
        check(rc)
    return out.raw[:96 * n], [status[i] for i in range(n)]


# ------------------------------------------------------------------------------------------------
# Device-resident columns and the polynomial side of the prover: everything below keeps
# its data in HBM between calls; only scalars and results the transcript needs cross the bus.
# ------------------------------------------------------------------------------------------------
class DeviceBuffer:
    """A raw HBM allocation made through the C ABI (no CUDA runtime on the caller's side)."""

    def __init__(self, nbytes: int):
        p = C.c_void_p(0)
        check(lib().b200zk_dev_alloc(C.byref(p), nbytes))
        self.ptr = p.value
        self.nbytes = nbytes

    def upload(self, data: bytes, offset: int = 0) -> None:
        assert offset + len(data) <= self.nbytes
        check(lib().b200zk_dev_upload(self.ptr + offset, addr(data), len(data)))

    def download(self, nbytes: Optional[int] = None, offset: int = 0) -> bytes:
        nbytes = self.nbytes - offset if nbytes is None else nbytes
        out = C.create_string_buffer(max(nbytes, 1))
        check(lib().b200zk_dev_download(addr(out), self.ptr + offset, nbytes))
        return out.raw[:nbytes]

    def free(self) -> None:
        if self.ptr:
            check(lib().b200zk_dev_free(self.ptr))
            self.ptr = None

    def __del__(self):   # best effort: columns that go out of scope give their HBM back
        try:
            if getattr(self, "ptr", None):
                lib().b200zk_dev_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class FrVec(DeviceBuffer):
    """n Fr elements in HBM, Montgomery form (the in-memory form of midnight_curves::Fq)."""

    def __init__(self, n: int):
        super().__init__(32 * max(n, 1))
        self.n = n

    @classmethod
    def from_canonical(cls, data: bytes) -> "FrVec":
        v = cls(len(data) // 32)
        if v.n:
            v.upload(data)
            check(lib().b200zk_fr_convert_dev(v.ptr, v.ptr, v.n, 1, None))
        return v

    @classmethod
    def from_ints(cls, xs: Sequence[int]) -> "FrVec":
        return cls.from_canonical(b"".join(fr_bytes(x) for x in xs))

    def to_canonical(self) -> bytes:
        if not self.n:
            return b""
        tmp = FrVec(self.n)
        check(lib().b200zk_fr_convert_dev(self.ptr, tmp.ptr, self.n, 0, None))
        out = tmp.download(32 * self.n)
        tmp.free()
        return out

    def to_ints(self) -> List[int]:
        b = self.to_canonical()
        return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(self.n)]


POINTWISE_MUL, POINTWISE_ADD, POINTWISE_SUB, POINTWISE_SCALE, POINTWISE_MULADD = 0, 1, 2, 3, 4


def fr_pointwise(op: int, a: FrVec, b: Optional[FrVec] = None, scalar: Optional[int] = None, out: Optional[FrVec] = None) -> FrVec:
    out = out or FrVec(a.n)
    sb = fr_bytes(scalar) if scalar is not None else None
    check(lib().b200zk_fr_pointwise_dev(op, a.ptr, b.ptr if b else None, addr(sb), out.ptr, a.n, None))
    return out


def fr_lincomb(polys: Sequence[FrVec], coeffs: Sequence[int], out: Optional[FrVec] = None) -> FrVec:
    """sum_k coeffs[k] * polys[k]: the x1 / x2 / x4 combinations of multi_open (pcs/kzg.rs:55-79)."""
    n = polys[0].n
    out = out or FrVec(n)
    ptrs = (C.c_void_p * len(polys))(*[p.ptr for p in polys])
    cb = b"".join(fr_bytes(c) for c in coeffs)
    check(lib().b200zk_fr_lincomb_dev(C.addressof(ptrs), addr(cb), len(polys), out.ptr, n, None))
    return out


def fr_batch_invert(v: FrVec, out: Optional[FrVec] = None) -> FrVec:
    out = out or FrVec(v.n)
    check(lib().b200zk_fr_batch_invert_dev(v.ptr, out.ptr, v.n, None))
    return out


def fr_running_product(v: FrVec, init: Optional[int] = None, inclusive: bool = False, out: Optional[FrVec] = None) -> FrVec:
    """z_0 = init, z_{i+1} = z_i v_i (permutation / lookup grand products)."""
    out = out or FrVec(v.n)
    ib = fr_bytes(init) if init is not None else None
    check(lib().b200zk_fr_running_product_dev(v.ptr, out.ptr, v.n, addr(ib), 1 if inclusive else 0, None))
    return out


def fr_kate_div(p: FrVec, z: int, want_quotient: bool = True) -> Tuple[Optional[FrVec], int]:
    """(q, p(z)) with p(X) - p(z) = q(X) (X - z), one sweep over the coefficients."""
    quot = FrVec(max(p.n - 1, 0)) if want_quotient else None
    ev = FrVec(1)
    zb = fr_bytes(z)
    check(lib().b200zk_fr_kate_div_dev(p.ptr, p.n, addr(zb), quot.ptr if quot else None, ev.ptr, None))
    e = ev.to_ints()[0]
    ev.free()
    return quot, e


class GateProgram:
    """A gate program resident on the device: the numerator of the quotient polynomial as a register
    machine over the extended-domain columns (include/b200zk.h, b200zk_gate_program_*)."""

    def __init__(self, words: Sequence[int], consts: Sequence[int], rotations: Sequence[int], n_columns: int, k: int,
                 extended_k: int, t_inv: Optional[Sequence[int]] = None):
        self.extended_k = extended_k
        self.n_columns = n_columns
        w = (C.c_uint32 * max(len(words), 1))(*words)
        cb = b"".join(fr_bytes(c) for c in consts)
        rot = (C.c_int32 * max(len(rotations), 1))(*rotations)
        tb, log_period = None, 0
        if t_inv is not None:
            tb = b"".join(fr_bytes(t) for t in t_inv)
            log_period = len(t_inv).bit_length() - 1
            assert 1 << log_period == len(t_inv)
        h = C.c_uint64(0)
        check(lib().b200zk_gate_program_create(C.addressof(w), len(words) // 4, addr(cb), len(consts), C.addressof(rot),
                                               len(rotations), addr(tb), log_period, n_columns, k, extended_k, C.byref(h)))
        self.handle = h.value

    def set_const(self, index: int, value: int) -> None:
        vb = fr_bytes(value)
        check(lib().b200zk_gate_program_set_const(self.handle, index, addr(vb)))

    def run(self, columns: Sequence[FrVec], out: Optional[FrVec] = None, accumulate: bool = False) -> FrVec:
        assert len(columns) == self.n_columns
        out = out or FrVec(1 << self.extended_k)
        ptrs = (C.c_void_p * max(len(columns), 1))(*[c.ptr for c in columns])
        check(lib().b200zk_gate_program_run_dev(self.handle, C.addressof(ptrs), out.ptr, 1 if accumulate else 0, None))
        return out

    def release(self) -> None:
        if self.handle:
            check(lib().b200zk_gate_program_release(self.handle))
            self.handle = None


def srs_generate(s: int, k: int, want_lagrange: bool = True) -> Tuple[DeviceBuffer, Optional[DeviceBuffer]]:
    """ParamsKZG::unsafe_setup's two G1 tables for secret s, left in HBM as packed Montgomery affine
    points (src/kzg_params.rs:33-80 caches what this produces)."""
    n = 1 << k
    g = DeviceBuffer(96 * n)
    gl = DeviceBuffer(96 * n) if want_lagrange else None
    sb = fr_bytes(s)
    wb = fr_bytes(pow(ROOT_OF_UNITY, 1 << (TWO_ADICITY - k), R_MOD))
    check(lib().b200zk_srs_generate_dev(addr(sb), k, addr(wb), g.ptr, gl.ptr if gl else None, None))
    return g, gl


def g1_export(buf: DeviceBuffer, n: int) -> bytes:
    out = C.create_string_buffer(96 * max(n, 1))
    check(lib().b200zk_g1_export_dev(buf.ptr, n, addr(out)))
    return out.raw[:96 * n]


def params_unsafe_setup(k: int, s: int) -> ParamsKZG:
    """SRS generated and registered without leaving the device."""
    g, gl = srs_generate(s, k)
    params = ParamsKZG.from_device(k, g.ptr, gl.ptr)
    g.free()
    gl.free()
    return params


# ------------------------------------------------------------------------------------------------
# Expression trees -> gate programs.  The Rust side does this once per circuit over halo2's
# ``Expression<F>`` with the same visitor the reference uses to walk gates
# (/root/reference/src/plutus_gen/extraction/mod.rs:81-102); this is the same compiler for the Python
# mirror and the tests.  An expression is a nested tuple:
#   ("const", int) | ("challenge", i) | ("query", column, rotation) |
#   ("neg", e) | ("sum", a, b) | ("product", a, b) | ("scaled", e, int)
# ------------------------------------------------------------------------------------------------
_OP = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "double": 4, "square": 5, "muladd": 6, "mov": 7}
GATE_MAX_REGS = 48


class CompiledGates:
    def __init__(self, words, consts, rotations, challenge_slots, n_columns):
        self.words, self.consts, self.rotations = words, consts, rotations
        self.challenge_slots = challenge_slots      # challenge index -> constant-table index (for set_const)
        self.n_columns = n_columns

    def instantiate(self, k: int, extended_k: int, t_inv: Optional[Sequence[int]] = None) -> "GateProgram":
        return GateProgram(self.words, self.consts, self.rotations, self.n_columns, k, extended_k, t_inv)


def compile_gates(gates: Sequence[tuple], n_columns: int, y_challenge: Optional[int] = None) -> CompiledGates:
    """Compiles the gate expressions into one program whose result is their fold sum_j gate_j * y^(len-1-j) (upstream
    folds the gates of the quotient numerator with the challenge y); with one gate and no y it is the gate itself."""
    consts: List[int] = []
    const_idx: dict = {}
    chal: dict = {}
    rots: List[int] = []
    rot_idx: dict = {}
    words: List[int] = []
    free = list(range(GATE_MAX_REGS - 1, 0, -1))       # register 0 is the accumulator of the y-fold

    def K(v):
        key = ("k", v % R_MOD)
        if key not in const_idx:
            const_idx[key] = len(consts)
            consts.append(v % R_MOD)
        return (0 << 28) | const_idx[key]

    def CH(i):
        if i not in chal:
            chal[i] = len(consts)
            consts.append(0)
        return (0 << 28) | chal[i]

    def ROT(r):
        if r not in rot_idx:
            rot_idx[r] = len(rots)
            rots.append(r)
        return rot_idx[r]

    def emit(op, dst, a, b=0, c=0):
        words.extend([_OP[op] | (dst << 8), a, b, c])

    def alloc():
        if not free:
            raise ValueError("expression needs more than %d registers; split it into several programs" % GATE_MAX_REGS)
        return free.pop()

    def release(src):
        if src >> 28 == 1 and (src & 0x0FFFFFFF) != 0:
            free.append(src & 0x0FFFFFFF)

    need_cache: dict = {}

    def need(e) -> int:
        """Registers needed to evaluate e when the hungrier operand goes first (Sethi-Ullman numbering)."""
        key = id(e)
        if key not in need_cache:
            kind = e[0]
            if kind in ("const", "challenge", "query"):
                v = 0
            elif kind in ("neg", "scaled"):
                v = max(need(e[1]), 1)
            else:
                a, b = need(e[1]), need(e[2])
                v = max(a, b) if a != b else a + 1
                v = max(v, 1)
            need_cache[key] = v
        return need_cache[key]

    def go2(x, y):
        """Evaluates two operands, the one that needs more registers first; returns their sources in (x, y) order."""
        if need(y) > need(x):
            sy = go(y)
            return go(x), sy
        sx = go(x)
        return sx, go(y)

    def go(e):
        """Returns a source operand holding the value of e (a leaf is used in place, anything else lands in a register)."""
        kind = e[0]
        if kind == "const":
            return K(e[1])
        if kind == "challenge":
            return CH(e[1])
        if kind == "query":
            if not 0 <= e[1] < n_columns:
                raise ValueError("query of column %d outside the %d columns" % (e[1], n_columns))
            return (2 << 28) | (e[1] << 12) | ROT(e[2])
        if kind == "neg":
            a = go(e[1])
            release(a)
            d = alloc()
            emit("neg", d, a)
            return (1 << 28) | d
        if kind == "scaled":
            a = go(e[1])
            release(a)
            d = alloc()
            emit("mul", d, a, K(e[2]))
            return (1 << 28) | d
        if kind in ("sum", "product"):
            # a*b + c in one instruction when a sum has a product on one side
            if kind == "sum" and e[1][0] == "product":
                if need(e[2]) > need(e[1]):
                    z = go(e[2])
                    x, y_ = go2(e[1][1], e[1][2])
                else:
                    x, y_ = go2(e[1][1], e[1][2])
                    z = go(e[2])
                for s_ in (x, y_, z):
                    release(s_)
                d = alloc()
                emit("muladd", d, x, y_, z)
                return (1 << 28) | d
            a, b = go2(e[1], e[2])
            release(a)
            release(b)
            d = alloc()
            emit("add" if kind == "sum" else "mul", d, a, b)
            return (1 << 28) | d
        raise ValueError("unknown expression node %r" % (kind,))

    if not gates:
        raise ValueError("no gates")
    ysrc = CH(y_challenge) if y_challenge is not None else None
    for j, g in enumerate(gates):
        v = go(g)
        if j == 0:
            emit("mov", 0, v)
        elif ysrc is None:
            emit("add", 0, (1 << 28) | 0, v)
        else:
            emit("muladd", 0, (1 << 28) | 0, ysrc, v)
        release(v)
    return CompiledGates(words, consts, rots, chal, n_columns)
