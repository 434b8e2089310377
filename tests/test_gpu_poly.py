"""Parity of the polynomial-side Fr kernels with the big-int oracle, through the
C ABI: pointwise ops, linear combinations, batched inversion, running products, evaluation + Kate division,
and gate programs over the extended domain; plus size-independent identities at 2^20..2^22."""
import random

import pytest

pytestmark = pytest.mark.gpu

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def rand_fr(rng, n, zeros=0.0):
    return [0 if rng.random() < zeros else rng.randrange(R) for _ in range(n)]


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 2047, 2048, 2049, 4096 + 17, 70001])
def test_running_product_and_batch_invert(gpu, pyref, n):
    rng = random.Random(n)
    H = gpu.host
    v = rand_fr(rng, n, zeros=0.02 if n > 40 else 0.0)
    dv = H.FrVec.from_ints(v)
    assert H.fr_batch_invert(dv).to_ints() == pyref.batch_invert(v)
    nz = [x or 1 for x in v]
    dn = H.FrVec.from_ints(nz)
    assert H.fr_running_product(dn).to_ints() == pyref.running_product(nz)
    init = rng.randrange(R)
    assert H.fr_running_product(dn, init=init, inclusive=True).to_ints() == pyref.running_product(nz, init, True)
    # in place
    H.fr_running_product(dn, out=dn)
    assert dn.to_ints() == pyref.running_product(nz)
    H.fr_batch_invert(dv, out=dv)
    assert dv.to_ints() == pyref.batch_invert(v)


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 2047, 2048, 2049, 3 * 2048 + 5, 50000])
def test_kate_division_and_evaluation(gpu, pyref, n):
    rng = random.Random(1000 + n)
    H = gpu.host
    c = rand_fr(rng, n)
    for z in (rng.randrange(R), 0, 1, R - 1):
        dq, ev = H.fr_kate_div(H.FrVec.from_ints(c), z)
        q, e = pyref.kate_div(c, z)
        assert ev == e == pyref.poly_eval(c, z)
        assert dq.to_ints() == q
        _, ev2 = H.fr_kate_div(H.FrVec.from_ints(c), z, want_quotient=False)
        assert ev2 == e


def test_pointwise_and_lincomb(gpu, pyref):
    rng = random.Random(7)
    H = gpu.host
    n = 5000
    a, b = rand_fr(rng, n), rand_fr(rng, n)
    da, db = H.FrVec.from_ints(a), H.FrVec.from_ints(b)
    assert H.fr_pointwise(H.POINTWISE_MUL, da, db).to_ints() == [x * y % R for x, y in zip(a, b)]
    assert H.fr_pointwise(H.POINTWISE_ADD, da, db).to_ints() == [(x + y) % R for x, y in zip(a, b)]
    assert H.fr_pointwise(H.POINTWISE_SUB, da, db).to_ints() == [(x - y) % R for x, y in zip(a, b)]
    s = rng.randrange(R)
    assert H.fr_pointwise(H.POINTWISE_SCALE, da, scalar=s).to_ints() == [x * s % R for x in a]
    acc = H.FrVec.from_ints(a)
    H.fr_pointwise(H.POINTWISE_MULADD, da, db, out=acc)
    assert acc.to_ints() == [(x * y + x) % R for x, y in zip(a, b)]
    for count in (1, 3, 16, 17, 40):
        polys = [rand_fr(rng, 700) for _ in range(count)]
        coeffs = rand_fr(rng, count)
        got = H.fr_lincomb([H.FrVec.from_ints(p) for p in polys], coeffs).to_ints()
        assert got == pyref.lincomb(polys, coeffs), count
    # canonical <-> Montgomery round trip keeps the bytes
    raw = b"".join(x.to_bytes(32, "little") for x in a)
    assert H.FrVec.from_canonical(raw).to_canonical() == raw


def random_program(rng, pyref, n_cols, n_rot, n_consts, n_instr, max_regs=48):
    prog, written = [], []
    for _ in range(n_instr):
        def src():
            kinds = ["col", "const"] + (["reg"] * 2 if written else [])
            k = rng.choice(kinds)
            if k == "col":
                return pyref.gate_col(rng.randrange(n_cols), rng.randrange(n_rot))
            if k == "const":
                return pyref.gate_const(rng.randrange(n_consts))
            return pyref.gate_reg(rng.choice(written))
        op = rng.randrange(8)
        dst = rng.randrange(max_regs)
        prog.append((op, dst, src(), src(), src()))
        if dst not in written:
            written.append(dst)
    return prog


@pytest.mark.parametrize("k,ext_k,seed", [(3, 5, 1), (4, 4, 2), (5, 7, 3), (6, 8, 4)])
def test_gate_program_vs_oracle(gpu, pyref, k, ext_k, seed):
    rng = random.Random(seed)
    H = gpu.host
    n_ext = 1 << ext_k
    n_cols, rotations = 5, [0, 1, -1, 2, -3]
    consts = rand_fr(rng, 6)
    cols = [rand_fr(rng, n_ext) for _ in range(n_cols)]
    prog = random_program(rng, pyref, n_cols, len(rotations), len(consts), 60)
    t_inv = pyref.vanishing_inverse_on_coset(7, k, ext_k) if ext_k > k else None
    gp = H.GateProgram(pyref.gate_program_words(prog), consts, rotations, n_cols, k, ext_k, t_inv)
    dcols = [H.FrVec.from_ints(c) for c in cols]
    out = gp.run(dcols)
    want = pyref.gate_eval(prog, consts, rotations, cols, k, ext_k, t_inv)
    assert out.to_ints() == want
    # accumulate mode and a changed challenge
    gp.set_const(2, 12345)
    consts2 = list(consts)
    consts2[2] = 12345
    gp.run(dcols, out=out, accumulate=True)
    want2 = pyref.gate_eval(prog, consts2, rotations, cols, k, ext_k, t_inv)
    assert out.to_ints() == [(a + b) % R for a, b in zip(want, want2)]
    gp.release()


def test_gate_program_plonk_gate_vanishes(gpu, pyref):
    """The standard arithmetic gate q_m a b + q_l a + q_r b + q_o c + q_c over a satisfying trace: the numerator
    vanishes on H, so the quotient times (X^n - 1) reproduces it on the coset, and the quotient has degree < 2n."""
    rng = random.Random(99)
    H = gpu.host
    k, ext_k = 6, 8
    n, n_ext = 1 << k, 1 << ext_k
    a, b = rand_fr(rng, n), rand_fr(rng, n)
    qm, ql, qr, qo = rand_fr(rng, n), rand_fr(rng, n), rand_fr(rng, n), [R - 1] * n
    c = [(qm[i] * a[i] * b[i] + ql[i] * a[i] + qr[i] * b[i]) % R for i in range(n)]
    qc = [0] * n
    dom = H.EvaluationDomain(4, k, g_coset=7)
    assert dom.extended_k == ext_k

    def extend(vals):  # Lagrange values -> extended coset evaluations, staying on the device after the upload
        coeff = dom.lagrange_to_coeff(b"".join(x.to_bytes(32, "little") for x in vals))
        return H.FrVec.from_canonical(dom.coeff_to_extended(coeff))

    cols = [extend(v) for v in (a, b, c, qm, ql, qr, qo, qc)]
    P = pyref
    prog = [
        (P.GATE_OPS["mul"], 0, P.gate_col(0, 0), P.gate_col(1, 0), 0),
        (P.GATE_OPS["mul"], 0, P.gate_reg(0), P.gate_col(3, 0), 0),
        (P.GATE_OPS["muladd"], 0, P.gate_col(4, 0), P.gate_col(0, 0), P.gate_reg(0)),
        (P.GATE_OPS["muladd"], 0, P.gate_col(5, 0), P.gate_col(1, 0), P.gate_reg(0)),
        (P.GATE_OPS["muladd"], 0, P.gate_col(6, 0), P.gate_col(2, 0), P.gate_reg(0)),
        (P.GATE_OPS["add"], 0, P.gate_reg(0), P.gate_col(7, 0), 0),
    ]
    t_inv = P.vanishing_inverse_on_coset(7, k, ext_k)
    gp = H.GateProgram(P.gate_program_words(prog), [], [0], 8, k, ext_k, t_inv)
    h_ext = gp.run(cols)
    h_coeff = dom.extended_to_coeff(h_ext.to_canonical())           # n * 3 low coefficients
    full = H.EvaluationDomain._ntt                                   # all 2^ext_k coefficients: the top n must be zero
    buf = bytearray(h_ext.to_canonical())
    full(buf, ext_k, dom.extended_omega_inv, gpu.NTT_INVERSE_SCALE | gpu.NTT_COSET_OUT, dom.g_coset_inv)
    assert bytes(buf[:len(h_coeff)]) == h_coeff
    assert bytes(buf[32 * 2 * n:]) == bytes(32 * (n_ext - 2 * n)), "quotient degree must stay below 2n for a degree-3 gate"
    assert any(buf[:32 * n])
    gp.release()


def test_gate_program_rejects_bad_programs(gpu, pyref):
    H = gpu.host
    P = pyref
    with pytest.raises(gpu.B200zkError):   # reads a register that was never written
        H.GateProgram(P.gate_program_words([(0, 0, P.gate_reg(3), P.gate_const(0), 0)]), [1], [0], 1, 3, 4)
    with pytest.raises(gpu.B200zkError):   # column out of range
        H.GateProgram(P.gate_program_words([(7, 0, P.gate_col(2, 0), 0, 0)]), [1], [0], 1, 3, 4)
    with pytest.raises(gpu.B200zkError):   # destination register out of range
        H.GateProgram(P.gate_program_words([(7, 48, P.gate_const(0), 0, 0)]), [1], [0], 1, 3, 4)


@pytest.mark.parametrize("log_n", [20, 22])
def test_large_identities(gpu, oracle, pyref, log_n):
    """Size-independent properties at sizes the big-int oracle cannot walk: (i) the quotient and the evaluation of
    one sweep satisfy p(z') = q(z') (z' - z) + p(z) at a second point; (ii) the running product of v times the
    running product of 1/v is all ones."""
    H = gpu.host
    n = 1 << log_n
    rng = random.Random(log_n)
    p = H.FrVec.from_canonical(oracle.synth_scalars(77, 0, n))
    z, z2 = rng.randrange(R), rng.randrange(R)
    q, pz = H.fr_kate_div(p, z)
    _, pz2 = H.fr_kate_div(p, z2, want_quotient=False)
    _, qz2 = H.fr_kate_div(q, z2, want_quotient=False)
    assert pz2 == (qz2 * (z2 - z) + pz) % R
    inv = H.fr_batch_invert(p)
    prod = H.fr_running_product(p, inclusive=True)
    iprod = H.fr_running_product(inv, inclusive=True)
    ones = H.fr_pointwise(H.POINTWISE_MUL, prod, iprod).to_canonical()
    assert ones == (1).to_bytes(32, "little") * n


def test_compiled_gates_on_device(gpu, pyref):
    """Expression trees -> host.compile_gates -> device program over extended-domain columns (rotations scaled by
    2^(ext_k - k), challenges set after creation), against direct evaluation of the trees."""
    from test_oracle_poly import _eval_expr, _sample_gates
    rng = random.Random(21)
    H = gpu.host
    k, ext_k = 5, 7
    n_ext, scale = 1 << ext_k, 1 << (ext_k - k)
    cols = [rand_fr(rng, n_ext) for _ in range(6)]
    challenges = {0: rng.randrange(R), 1: rng.randrange(R), 2: rng.randrange(R)}
    gates = _sample_gates()
    cg = H.compile_gates(gates, 6, y_challenge=0)
    gp = cg.instantiate(k, ext_k)
    for ci, slot in cg.challenge_slots.items():
        gp.set_const(slot, challenges[ci])
    out = gp.run([H.FrVec.from_ints(c) for c in cols]).to_ints()

    def scaled(e):   # the same trees with rotations expressed in extended-domain rows
        if e[0] == "query":
            return ("query", e[1], e[2] * scale)
        if e[0] in ("const", "challenge"):
            return e
        if e[0] == "scaled":
            return ("scaled", scaled(e[1]), e[2])
        return (e[0],) + tuple(scaled(x) for x in e[1:])

    for row in range(n_ext):
        want = 0
        for g in gates:
            want = (want * challenges[0] + _eval_expr(scaled(g), cols, row, n_ext, challenges)) % R
        assert out[row] == want, row
    gp.release()


def test_extend_and_power_table(gpu, pyref):
    """b200zk_fr_extend_dev (zero padding of coeff_to_extended on the device, batched) and b200zk_fr_power_table_dev
    (the omega^(i2 k1) block of the multi-GPU transform) against their definitions."""
    import ctypes as C
    H = gpu.host
    lib, chk = gpu.lib(), gpu.capi.check
    rng = random.Random(5)
    n_in, n_out, batch = 24, 64, 3
    vals = rand_fr(rng, n_in * batch)
    src = H.FrVec.from_ints(vals)
    dst = H.FrVec(n_out * batch)
    chk(lib.b200zk_fr_extend_dev(src.ptr, n_in, dst.ptr, n_out, batch, None))
    got = dst.to_ints()
    for b in range(batch):
        assert got[b * n_out:b * n_out + n_in] == vals[b * n_in:(b + 1) * n_in]
        assert got[b * n_out + n_in:(b + 1) * n_out] == [0] * (n_out - n_in)
    base = rng.randrange(R)
    row0, rows, cols = 5, 7, 9
    tab = H.FrVec(rows * cols)
    bb = base.to_bytes(32, "little")
    chk(lib.b200zk_fr_power_table_dev(gpu.capi.addr(bb), row0, rows, cols, tab.ptr, None))
    assert tab.to_ints() == [pow(base, (row0 + r) * c, R) for r in range(rows) for c in range(cols)]


def test_full_size_parity_vs_c_oracle(gpu, oracle):
    """Bit-exact at 2^20 (every element, not a property): Kate division + evaluation, running product, batched inversion
    and an 8-term linear combination against the C restatements of the checker."""
    import hashlib
    H = gpu.host
    n = 1 << 20
    raw = oracle.synth_scalars(91, 0, n)
    v = H.FrVec.from_canonical(raw)
    z = (0x1234567 ** 9 % R).to_bytes(32, "little")
    q, e = H.fr_kate_div(v, int.from_bytes(z, "little"))
    wq, we = oracle.kate_div(raw, z)
    assert e == int.from_bytes(we, "little")
    assert hashlib.sha256(q.to_canonical()).digest() == hashlib.sha256(wq).digest()
    assert H.fr_running_product(v).to_canonical() == oracle.running_product(raw)
    init = (987654321).to_bytes(32, "little")
    assert H.fr_running_product(v, init=987654321, inclusive=True).to_canonical() == oracle.running_product(raw, init, True)
    holes = bytearray(raw)
    for i in (0, 5, n // 2, n - 1):
        holes[32 * i:32 * i + 32] = bytes(32)
    assert H.fr_batch_invert(H.FrVec.from_canonical(bytes(holes))).to_canonical() == oracle.batch_invert(bytes(holes))
    m = 1 << 16
    polys = [oracle.synth_scalars(200 + k, 0, m) for k in range(8)]
    coeffs = [(k + 3) ** 40 % R for k in range(8)]
    got = H.fr_lincomb([H.FrVec.from_canonical(p) for p in polys], coeffs).to_canonical()
    assert got == oracle.lincomb(b"".join(polys), b"".join(c.to_bytes(32, "little") for c in coeffs), 8)
