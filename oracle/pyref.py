"""Pure-Python big-int restatement of the BLS12-381 G1 MSM / Fr NTT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
(`plutus-halo2-verifier-gen_b200/`); only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may use it, and only
as the checker.

Parity status: the arithmetic of this path lives in crates that are NOT under
/root/reference (midnight-proofs =0.8.0, midnight-curves =0.3.0 -> blst; see
/root/reference/Cargo.toml:27-28), so the reference binary cannot be run here.
This file restates the published algorithms (BLS12-381 per the ZCash spec,
halo2 `best_fft` DFT convention) and is PINNED against every known-answer
vector the reference's own Aiken / Plinth tests hold for this path
(tests/test_oracle_kats.py).  What no reference test pins (an MSM of size > 1,
an NTT output vector) is "parity unpinned" there and rests on mathematical
uniqueness of the result plus the pinned encodings.

Reference anchors (relative to /root/reference):
  moduli            plinth-verifier/plutus-halo2/src/Plutus/Crypto/BlsTypes.hs:97,102-103
                    aiken-verifier/aiken_halo2/lib/bls_utils.ak:14-15
  generator 7/DELTA plinth-verifier/plutus-halo2/src/Plutus/Crypto/Constants.hs:10-13
  omega_k           aiken-verifier/aiken_halo2/lib/omega_rotations.ak:48-81
  G1 encoding       aiken-verifier/aiken_halo2/lib/bls_utils.ak:17-49,
                    plinth-verifier/.../Halo2/CompressUncompress.hs:70-100
  Fr encoding       aiken-verifier/aiken_halo2/lib/transcript.ak:29-45,158-179
  transcript        src/plutus_gen/adjusted_types/mod.rs:30-72,
                    aiken-verifier/aiken_halo2/lib/transcript.ak:85-106
  lagrange basis    aiken-verifier/aiken_halo2/lib/lagrange.ak:80-100
  H2MO verifier     plinth-verifier/.../Halo2/Halo2MultiOpenMSM.hs:60-189
"""
from __future__ import annotations

import hashlib

# --------------------------------------------------------------------------- fields
# BlsTypes.hs:97 (scalar field, called Fq/BlsScalar in midnight-curves)
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
# BlsTypes.hs:102-103 / bls_utils.ak:14-15 (base field Fp)
P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
CURVE_B = 4
# standard generator of G1 (its compressed form is the KAT at transcript.ak:125)
G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G1_GEN = (G1_X, G1_Y)
INF = None  # point at infinity; on the wire it is affine (0,0) (CompressUncompress.hs:72)

MULT_GEN = 7  # Constants.hs:10-13: DELTA = 7^(2^32)
TWO_ADICITY = 32
ROOT_OF_UNITY = pow(MULT_GEN, (R_MOD - 1) >> TWO_ADICITY, R_MOD)
DELTA = pow(MULT_GEN, 1 << TWO_ADICITY, R_MOD)


def omega(k: int) -> int:
    """2^k-th primitive root of unity used by the halo2 EvaluationDomain
    (convention pinned by omega_rotations.ak:50-52 for k = 14)."""
    assert 0 <= k <= TWO_ADICITY
    return pow(ROOT_OF_UNITY, 1 << (TWO_ADICITY - k), R_MOD)


def fr_inv(a: int) -> int:
    return pow(a, R_MOD - 2, R_MOD)


def fr_from_le(b: bytes) -> int:
    """32-byte little-endian, reduced mod r on read (transcript.ak:158-179)."""
    assert len(b) == 32
    return int.from_bytes(b, "little") % R_MOD


def fr_to_le(a: int) -> bytes:
    return (a % R_MOD).to_bytes(32, "little")


# --------------------------------------------------------------------------- G1
def g1_is_on_curve(pt) -> bool:
    if pt is INF:
        return True
    x, y = pt
    return (y * y - (x * x * x + CURVE_B)) % P_MOD == 0


def g1_neg(pt):
    if pt is INF:
        return INF
    return (pt[0], (-pt[1]) % P_MOD)


def g1_add(a, b):
    if a is INF:
        return b
    if b is INF:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return INF
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    y3 = (lam * (x1 - x3) - y1) % P_MOD
    return (x3, y3)


def g1_mul(pt, k: int):
    k %= R_MOD
    acc = INF
    add = pt
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


def g1_msm_naive(points, scalars):
    acc = INF
    for pt, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(pt, s))
    return acc


def fp_sqrt(a: int):
    """p = 3 mod 4 -> a^((p+1)/4) (CompressUncompress.hs:98)."""
    s = pow(a, (P_MOD + 1) // 4, P_MOD)
    return s if s * s % P_MOD == a % P_MOD else None


def g1_compress(pt) -> bytes:
    """ZCash 48-byte big-endian compressed encoding; flag bits 0x80 compressed,
    0x40 infinity, 0x20 'y is the lexicographically larger root'
    (bls_utils.ak:17-49)."""
    if pt is INF:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    flags = 0x80
    if y > (P_MOD - y) % P_MOD:
        flags |= 0x20
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= flags
    return bytes(b)


def g1_decompress(b: bytes):
    assert len(b) == 48
    flags = b[0] & 0xE0
    if not flags & 0x80:
        raise ValueError("not a compressed encoding")
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if flags & 0x40:
        if x != 0 or flags & 0x20:
            raise ValueError("bad infinity encoding")
        return INF
    if x >= P_MOD:
        raise ValueError("x out of range")
    y = fp_sqrt((x * x * x + CURVE_B) % P_MOD)
    if y is None:
        raise ValueError("not on curve")
    if (y > P_MOD - y) != bool(flags & 0x20):
        y = P_MOD - y
    return (x, y)


def g1_to_wire(pt) -> bytes:
    """C-ABI affine wire format: x||y, 48-byte little-endian canonical each,
    identity = (0,0)."""
    if pt is INF:
        return bytes(96)
    return pt[0].to_bytes(48, "little") + pt[1].to_bytes(48, "little")


def g1_from_wire(b: bytes):
    assert len(b) == 96
    x = int.from_bytes(b[:48], "little")
    y = int.from_bytes(b[48:], "little")
    if x == 0 and y == 0:
        return INF
    return (x, y)


# --------------------------------------------------------------------------- NTT
def ntt_naive(a, w):
    """O(n^2) DFT: X[k] = sum_i a[i] w^(ik)  (halo2 best_fft convention,
    natural order in and out)."""
    n = len(a)
    out = []
    for k in range(n):
        wk = pow(w, k, R_MOD)
        acc = 0
        x = 1
        for i in range(n):
            acc = (acc + a[i] * x) % R_MOD
            x = x * wk % R_MOD
        out.append(acc)
    return out


def ntt(a, w):
    """Iterative radix-2 DIT with bit-reversal, same result as ntt_naive."""
    n = len(a)
    if n == 1:
        return list(a)
    logn = n.bit_length() - 1
    assert 1 << logn == n
    a = list(a)
    for i in range(n):
        j = int(format(i, "0%db" % logn)[::-1], 2)
        if i < j:
            a[i], a[j] = a[j], a[i]
    m = 1
    while m < n:
        wm = pow(w, n // (2 * m), R_MOD)
        for s in range(0, n, 2 * m):
            x = 1
            for j in range(m):
                t = a[s + j + m] * x % R_MOD
                u = a[s + j]
                a[s + j] = (u + t) % R_MOD
                a[s + j + m] = (u - t) % R_MOD
                x = x * wm % R_MOD
        m *= 2
    return a


def intt(a, w):
    n = len(a)
    ninv = fr_inv(n)
    return [x * ninv % R_MOD for x in ntt(a, fr_inv(w))]


def coset_ntt(a, w, g):
    """Evaluate on the coset g*H: scale a[i] by g^i, then DFT."""
    out, x = [], 1
    for v in a:
        out.append(v * x % R_MOD)
        x = x * g % R_MOD
    return ntt(out, w)


def coset_intt(a, w, g):
    """Inverse of coset_ntt: inverse DFT then scale by g^-i."""
    c = intt(a, w)
    gi = fr_inv(g)
    out, x = [], 1
    for v in c:
        out.append(v * x % R_MOD)
        x = x * gi % R_MOD
    return out


# --------------------------------------------------------------------------- transcript
class Transcript:
    """CardanoFriendlyBlake2b: unkeyed blake2b-256 over the whole absorbed
    history; prefix 0x01 per absorbed item, 0x00 per squeeze; challenge =
    LE(H) + LE(H(H)) * 2^256 mod r  (adjusted_types/mod.rs:39-60,
    transcript.ak:85-106)."""

    R256 = (1 << 256) % R_MOD  # transcript.ak:99 / Transcript.hs:78-79

    def __init__(self, proof: bytes = b""):
        self.buf = bytearray()
        self.proof = bytes(proof)
        self.pos = 0

    def common_scalar(self, s: int):
        self.buf += b"\x01" + fr_to_le(s)

    def common_point_bytes(self, b: bytes):
        assert len(b) == 48
        self.buf += b"\x01" + b

    def common_point(self, pt):
        self.common_point_bytes(g1_compress(pt))

    def read_scalar(self) -> int:
        b = self.proof[self.pos:self.pos + 32]
        self.pos += 32
        self.buf += b"\x01" + b
        return fr_from_le(b)

    def read_point_bytes(self) -> bytes:
        b = self.proof[self.pos:self.pos + 48]
        self.pos += 48
        self.buf += b"\x01" + b
        return b

    def read_point(self):
        return g1_decompress(self.read_point_bytes())

    def squeeze(self) -> int:
        self.buf += b"\x00"
        h = hashlib.blake2b(bytes(self.buf), digest_size=32).digest()
        hh = hashlib.blake2b(h, digest_size=32).digest()
        return (int.from_bytes(h, "little") + int.from_bytes(hh, "little") * self.R256) % R_MOD


# --------------------------------------------------------------------------- verifier scalar side
def lagrange_basis(x: int, xn: int, barycentric_weight: int, rotations):
    """l_i(x) = (x^n - 1)/n * w_i / (x - w_i)   (lagrange.ak:80-100)."""
    common = (xn - 1) * barycentric_weight % R_MOD
    return [fr_inv((x - w) % R_MOD) * common % R_MOD * w % R_MOD for w in rotations]


def lagrange_evaluation(points, x: int) -> int:
    """Evaluate at x the polynomial through (xi, yi) (lagrange.ak:40-78)."""
    acc = 0
    for xi, yi in points:
        num, den = 1, 1
        for xj, _ in points:
            if xj != xi:
                num = num * (x - xj) % R_MOD
                den = den * (xi - xj) % R_MOD
        acc = (acc + yi * num % R_MOD * fr_inv(den)) % R_MOD
    return acc


def h2mo_q_eval_sets(commitment_map, n_sets: int, x1: int):
    """q_eval_sets[s][p] = sum_j x1^j * eval_{s,j}[p] (Halo2MultiOpenMSM.hs:149-189).
    commitment_map entries: (commitment, set_index, points, evals)."""
    out = []
    for s in range(n_sets):
        acc = None
        xp = 1
        for (_c, si, _pts, evals) in commitment_map:
            if si != s:
                continue
            scaled = [e * xp % R_MOD for e in evals]
            acc = scaled if acc is None else [(a + b) % R_MOD for a, b in zip(acc, scaled)]
            xp = xp * x1 % R_MOD
        out.append(acc or [])
    return out


def h2mo_f_eval(point_sets, q_eval_sets, x2: int, x3: int, proof_q_evals) -> int:
    """fold_{x2}((q_eval_s - r_s(x3)) / prod_{p in set s}(x3 - p))
    (Halo2MultiOpenMSM.hs:126-145; the fold runs over the reversed list)."""
    acc = 0
    for pts, evals, pq in reversed(list(zip(point_sets, q_eval_sets, proof_q_evals))):
        r_eval = lagrange_evaluation(list(zip(pts, evals)), x3)
        den = 1
        for p_ in pts:
            den = den * (x3 - p_) % R_MOD
        acc = (acc * x2 + (pq - r_eval) * fr_inv(den)) % R_MOD
    return acc


def h2mo_v(f_eval: int, x4: int, proof_q_evals) -> int:
    """v = sum_s x4^s q_eval_s + x4^S f_eval (Halo2MultiOpenMSM.hs:100-109)."""
    acc, xp = 0, 1
    for e in list(proof_q_evals) + [f_eval]:
        acc = (acc + xp * e) % R_MOD
        xp = xp * x4 % R_MOD
    return acc


def h2mo_sets(queries):
    """precompute_intermediate_sets (src/plutus_gen/extraction/pcs/mod.rs:36-109) on (polynomial, point[, eval]) queries:
    polynomials in order of first appearance, equal point sets share an index numbered by first appearance.
    Returns (point_sets, members, evals_by_poly) with the points of a set in ascending order."""
    per, order = {}, []
    for q in queries:
        p, x = q[0], q[1] % R_MOD
        if p not in per:
            per[p] = {}
            order.append(p)
        if len(q) > 2:
            per[p][x] = q[2] % R_MOD
        else:
            per[p].setdefault(x, None)
    point_sets, members, evals = [], [], {}
    for p in order:
        key = sorted(per[p])
        if key not in point_sets:
            point_sets.append(key)
            members.append([])
        members[point_sets.index(key)].append(p)
        evals[p] = [per[p][x] for x in key]
    return point_sets, members, evals


def poly_interpolate(points):
    """Coefficients (low to high) of the polynomial through (xi, yi)."""
    coef = [0] * len(points)
    for i, (xi, yi) in enumerate(points):
        num, den = [1], 1
        for j, (xj, _) in enumerate(points):
            if j != i:
                num = [(a - xj * b) % R_MOD for a, b in zip([0] + num, num + [0])]
                den = den * (xi - xj) % R_MOD
        sc = yi * fr_inv(den) % R_MOD
        for d, c in enumerate(num):
            coef[d] = (coef[d] + sc * c) % R_MOD
    return coef


def h2mo_open(polys, queries, transcript: "Transcript", commit):
    """Prover side of the halo2 multi-open in the message order of src/plutus_gen/extraction/pcs/kzg.rs:55-79 (X1, X2,
    FCommitment, X3, QEvals, X4, PI), on coefficient lists with big integers.  `commit` maps a coefficient list to a G1
    point.  Returns the proof bytes f || q_evals || pi."""
    n = len(polys[0])
    point_sets, members, _ = h2mo_sets(queries)
    x1 = transcript.squeeze()
    qs = []
    for mem in members:
        acc, xp = [0] * n, 1
        for p in mem:
            acc = [(a + xp * c) % R_MOD for a, c in zip(acc, polys[p])]
            xp = xp * x1 % R_MOD
        qs.append(acc)
    x2 = transcript.squeeze()
    f, xp = [0] * n, 1
    for pts, q in zip(point_sets, qs):
        r = poly_interpolate([(x, poly_eval(q, x)) for x in pts])
        g = [(c - (r[i] if i < len(r) else 0)) % R_MOD for i, c in enumerate(q)]
        for x in pts:
            g, rem = kate_div(g, x)
            assert rem == 0
        f = [(a + xp * (g[i] if i < len(g) else 0)) % R_MOD for i, a in enumerate(f)]
        xp = xp * x2 % R_MOD
    f_pt = commit(f)
    transcript.common_point(f_pt)
    x3 = transcript.squeeze()
    q_evals = [poly_eval(q, x3) for q in qs]
    for e in q_evals:
        transcript.common_scalar(e)
    x4 = transcript.squeeze()
    final, xp = [0] * n, 1
    for q in qs + [f]:
        final = [(a + xp * c) % R_MOD for a, c in zip(final, q)]
        xp = xp * x4 % R_MOD
    w, _v = kate_div(final, x3)
    pi_pt = commit(w)
    transcript.common_point(pi_pt)
    return g1_compress(f_pt) + b"".join(fr_to_le(e) for e in q_evals) + g1_compress(pi_pt)


# --------------------------------------------------------------------------- synthetic inputs
def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def synth_base_dlog(seed: int, i: int) -> int:
    """Discrete log a_i of synthetic base i: P_i = a_i * G, a_i = SplitMix64(seed + i)
    (never 0 for the seeds used)."""
    return splitmix64((seed + i) & 0xFFFFFFFFFFFFFFFF)


# --------------------------------------------------------------------------- polynomial side
# Checker for the Fr vector kernels next to the hot path.  Values are Python ints mod r.
def poly_eval(coeffs, z: int) -> int:
    """p(z) by Horner's rule (the evals the prover writes, extraction_steps/proof.rs:82-143)."""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % R_MOD
    return acc


def kate_div(coeffs, z: int):
    """Synthetic division: (q, p(z)) with p(X) - p(z) = q(X) (X - z); the witness polynomials of the
    KZG multi-open (src/plutus_gen/extraction/pcs/kzg.rs:55-79; verifier twin halo2_kzg.ak:46-171)."""
    n = len(coeffs)
    if n == 0:
        return [], 0
    q = [0] * (n - 1)
    s = 0
    for i in range(n - 1, -1, -1):
        s = (s * z + coeffs[i]) % R_MOD
        if i:
            q[i - 1] = s
    return q, s


def running_product(v, init: int = 1, inclusive: bool = False):
    """z_0 = init, z_{i+1} = z_i * v_i: the grand-product columns of the permutation and lookup
    arguments (extraction_steps/permutation.rs)."""
    out, acc = [], init % R_MOD
    for x in v:
        nxt = acc * x % R_MOD
        out.append(nxt if inclusive else acc)
        acc = nxt
    return out


def batch_invert(v):
    """Elementwise inverse; zeros stay zero (halo2's batch_invert convention)."""
    return [fr_inv(x) if x % R_MOD else 0 for x in v]


def lincomb(polys, coeffs):
    n = len(polys[0]) if polys else 0
    return [sum(c * p[i] for c, p in zip(coeffs, polys)) % R_MOD for i in range(n)]


def vanishing_inverse_on_coset(g_coset: int, k: int, ext_k: int):
    """1 / (X^n - 1) at X = g * w_ext^j: only 2^(ext_k - k) distinct values, because w_ext^n has that
    order (the division by the vanishing polynomial of the quotient, docs 'Vanishing')."""
    n = 1 << k
    w_ext = omega(ext_k)
    return [fr_inv((pow(g_coset * pow(w_ext, j, R_MOD) % R_MOD, n, R_MOD) - 1) % R_MOD) for j in range(1 << (ext_k - k))]


GATE_OPS = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "double": 4, "square": 5, "muladd": 6, "mov": 7}


def gate_const(i):
    return (0 << 28) | i


def gate_reg(i):
    return (1 << 28) | i


def gate_col(col, rot_idx):
    return (2 << 28) | (col << 12) | rot_idx


def gate_eval(program, consts, rotations, columns, log_n: int, log_ext: int, t_inv=None):
    """Reference evaluation of a gate program (include/b200zk.h, b200zk_gate_program_*) at every row
    of the extended domain: the numerator of the quotient h(X) walked gate by gate as the reference
    walks the same expressions (src/plutus_gen/extraction/mod.rs:81-102), then divided by X^n - 1.
    program: list of (op, dst, a, b, c) with encoded sources."""
    n_ext = 1 << log_ext
    scale = 1 << (log_ext - log_n)
    out = []
    for row in range(n_ext):
        regs = {}

        def src(s):
            kind, pay = s >> 28, s & 0x0FFFFFFF
            if kind == 0:
                return consts[pay] % R_MOD
            if kind == 1:
                return regs[pay]
            return columns[pay >> 12][(row + rotations[pay & 0xFFF] * scale) % n_ext]

        last = None
        for (op, dst, a, b, c) in program:
            x = src(a)
            if op == 0:
                r = x + src(b)
            elif op == 1:
                r = x - src(b)
            elif op == 2:
                r = x * src(b)
            elif op == 3:
                r = -x
            elif op == 4:
                r = 2 * x
            elif op == 5:
                r = x * x
            elif op == 6:
                r = x * src(b) + src(c)
            else:
                r = x
            regs[dst] = r % R_MOD
            last = dst
        v = regs[last] if last is not None else 0
        if t_inv is not None:
            v = v * t_inv[row % len(t_inv)] % R_MOD
        out.append(v)
    return out


def gate_program_words(program):
    """4 x u32 per instruction, the C-ABI encoding."""
    words = []
    for (op, dst, a, b, c) in program:
        words += [op | (dst << 8), a, b, c]
    return words


# --------------------------------------------------------------------------- SRS
def srs_scalars(s: int, k: int):
    """Discrete logs of ParamsKZG::unsafe_setup's two tables (src/kzg_params.rs:33-80 caches them):
    g[i] = s^i G;  g_lagrange[i] = L_i(s) G with L_i(s) = w^i (s^n - 1) / (n (s - w^i))."""
    n = 1 << k
    w = omega(k)
    mono = [pow(s, i, R_MOD) for i in range(n)]
    c = (pow(s, n, R_MOD) - 1) * fr_inv(n) % R_MOD
    lag = []
    wi = 1
    for _ in range(n):
        lag.append(wi * c % R_MOD * fr_inv((s - wi) % R_MOD) % R_MOD)
        wi = wi * w % R_MOD
    return mono, lag
