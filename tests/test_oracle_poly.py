"""CPU checks of the polynomial-side / SRS oracle functions (oracle/pyref.py) against independent formulations,
so that the GPU parity tests compare against a checker that has itself been cross-checked."""
import random

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def test_kate_div_reconstructs(pyref):
    rng = random.Random(1)
    for n in (1, 2, 9, 64):
        c = [rng.randrange(R) for _ in range(n)]
        z = rng.randrange(R)
        q, e = pyref.kate_div(c, z)
        assert e == pyref.poly_eval(c, z) == sum(ci * pow(z, i, R) for i, ci in enumerate(c)) % R
        rec = [0] * n
        for i, qi in enumerate(q):
            rec[i + 1] = (rec[i + 1] + qi) % R
            rec[i] = (rec[i] - qi * z) % R
        rec[0] = (rec[0] + e) % R
        assert rec == c


def test_running_product_and_inverse(pyref):
    rng = random.Random(2)
    v = [rng.randrange(1, R) for _ in range(50)] + [0]
    inv = pyref.batch_invert(v)
    assert inv[-1] == 0 and all(a * b % R == 1 for a, b in zip(v[:-1], inv[:-1]))
    ex = pyref.running_product(v[:-1], 5)
    assert ex[0] == 5 and ex[3] == 5 * v[0] * v[1] * v[2] % R
    assert pyref.running_product(v[:-1], 5, True)[:-1] == ex[1:]


def test_srs_scalars_are_the_lagrange_basis(pyref):
    s, k = 0xDEADBEEF, 4
    mono, lag = pyref.srs_scalars(s, k)
    n, w = 1 << k, pyref.omega(k)
    assert sum(lag) % R == 1
    rng = random.Random(3)
    f = [rng.randrange(R) for _ in range(n)]
    evals = pyref.ntt(f, w)
    assert sum(l * e for l, e in zip(lag, evals)) % R == sum(m * c for m, c in zip(mono, f)) % R == pyref.poly_eval(f, s)
    # same formula the verifier uses for its Lagrange evaluations (lagrange.ak:80-100)
    rot = [pow(w, i, R) for i in range(n)]
    assert lag == pyref.lagrange_basis(s, pow(s, n, R), pyref.fr_inv(n), rot)


def test_vanishing_inverse_periodicity(pyref):
    k, ext_k, g = 3, 5, 7
    t = pyref.vanishing_inverse_on_coset(g, k, ext_k)
    w = pyref.omega(ext_k)
    for j in range(1 << ext_k):
        x = g * pow(w, j, R) % R
        assert t[j % len(t)] * (pow(x, 1 << k, R) - 1) % R == 1


def test_gate_eval_matches_direct_formula(pyref):
    rng = random.Random(4)
    k = ext_k = 3
    n = 1 << k
    a, b = [rng.randrange(R) for _ in range(n)], [rng.randrange(R) for _ in range(n)]
    P = pyref
    prog = [(P.GATE_OPS["mul"], 1, P.gate_col(0, 0), P.gate_col(1, 1), 0),
            (P.GATE_OPS["muladd"], 2, P.gate_reg(1), P.gate_const(0), P.gate_col(0, 2)),
            (P.GATE_OPS["sub"], 0, P.gate_reg(2), P.gate_reg(1), 0)]
    got = P.gate_eval(prog, [5], [0, 1, -1], [a, b], k, ext_k)
    for i in range(n):
        m = a[i] * b[(i + 1) % n] % R
        assert got[i] == (m * 5 + a[(i - 1) % n] - m) % R
    assert P.gate_program_words(prog)[:4] == [2 | (1 << 8), 2 << 28, (2 << 28) | (1 << 12) | 1, 0]


def _eval_expr(e, cols, row, n, challenges):
    k = e[0]
    if k == "const":
        return e[1] % R
    if k == "challenge":
        return challenges[e[1]]
    if k == "query":
        return cols[e[1]][(row + e[2]) % n]
    if k == "neg":
        return -_eval_expr(e[1], cols, row, n, challenges) % R
    if k == "scaled":
        return _eval_expr(e[1], cols, row, n, challenges) * e[2] % R
    a, b = _eval_expr(e[1], cols, row, n, challenges), _eval_expr(e[2], cols, row, n, challenges)
    return (a + b) % R if k == "sum" else a * b % R


def _sample_gates():
    q = lambda c, r=0: ("query", c, r)
    plonk = ("sum", ("sum", ("product", ("product", q(3), q(0)), q(1)), ("product", q(4), q(0))),
             ("sum", ("product", q(5), q(1)), ("sum", ("neg", q(2)), ("const", 7))))
    perm = ("product", ("sum", q(0, 1), ("product", ("challenge", 1), q(1, -1))),
            ("sum", ("scaled", q(2), 5), ("challenge", 2)))
    boolean = ("product", q(0), ("sum", q(0), ("neg", ("const", 1))))
    return [plonk, perm, boolean]


def test_compile_gates_matches_tree_evaluation(zk, pyref):
    """host.compile_gates (the Python twin of the Rust-side compiler) against direct evaluation of the expression trees,
    through the checker's gate-program interpreter: single gates and the y-fold."""
    rng = random.Random(11)
    k = 4
    n = 1 << k
    cols = [[rng.randrange(R) for _ in range(n)] for _ in range(6)]
    challenges = {0: rng.randrange(R), 1: rng.randrange(R), 2: rng.randrange(R)}
    gates = _sample_gates()
    for subset, y in ((gates[:1], None), (gates, 0), (gates[1:], None)):
        cg = zk.host.compile_gates(subset, 6, y_challenge=y)
        consts = list(cg.consts)
        for ci, slot in cg.challenge_slots.items():
            consts[slot] = challenges[ci]
        prog = [(cg.words[4 * i] & 0xFF, cg.words[4 * i] >> 8, cg.words[4 * i + 1], cg.words[4 * i + 2], cg.words[4 * i + 3])
                for i in range(len(cg.words) // 4)]
        got = pyref.gate_eval(prog, consts, cg.rotations, cols, k, k)
        for row in range(n):
            vals = [_eval_expr(g, cols, row, n, challenges) for g in subset]
            if y is None:
                want = sum(vals) % R
            else:
                want = 0
                for v in vals:
                    want = (want * challenges[y] + v) % R
            assert got[row] == want, (row, y)


def test_c_oracle_polynomial_side_matches_bigint_oracle(oracle, pyref):
    """the C restatements used as the checker at 2^20 (oracle/b200zk_oracle.c) against the big-int ones"""
    rng = random.Random(8)
    tb = lambda v: b"".join(x.to_bytes(32, "little") for x in v)
    ti = lambda b: [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(len(b) // 32)]
    for n in (1, 2, 33, 500):
        v = [rng.randrange(R) for _ in range(n)]
        z = rng.randrange(R)
        q, e = oracle.kate_div(tb(v), z.to_bytes(32, "little"))
        wq, we = pyref.kate_div(v, z)
        assert ti(q) == wq and int.from_bytes(e, "little") == we
        assert ti(oracle.running_product(tb(v))) == pyref.running_product(v)
        init = rng.randrange(R)
        assert ti(oracle.running_product(tb(v), init.to_bytes(32, "little"), True)) == pyref.running_product(v, init, True)
        vz = list(v)
        vz[0] = 0
        assert ti(oracle.batch_invert(tb(vz))) == pyref.batch_invert(vz)
    polys = [[rng.randrange(R) for _ in range(40)] for _ in range(5)]
    cs = [rng.randrange(R) for _ in range(5)]
    assert ti(oracle.lincomb(b"".join(tb(p) for p in polys), tb(cs), 5)) == pyref.lincomb(polys, cs)


def test_compile_gates_random_trees(zk, pyref):
    """Random expression trees (hypothesis): the compiled register program evaluates to the tree's value at every row,
    never reads a register before writing it, and stays inside the register file."""
    from hypothesis import given, settings, strategies as st

    leaf = st.one_of(
        st.tuples(st.just("const"), st.integers(0, R - 1)),
        st.tuples(st.just("challenge"), st.integers(0, 2)),
        st.tuples(st.just("query"), st.integers(0, 3), st.integers(-2, 2)),
    )
    tree = st.recursive(
        leaf,
        lambda ch: st.one_of(
            st.tuples(st.just("neg"), ch),
            st.tuples(st.just("scaled"), ch, st.integers(0, R - 1)),
            st.tuples(st.just("sum"), ch, ch),
            st.tuples(st.just("product"), ch, ch),
        ),
        max_leaves=24,
    )
    rng = random.Random(17)
    k = 3
    n = 1 << k
    cols = [[rng.randrange(R) for _ in range(n)] for _ in range(4)]
    challenges = {0: rng.randrange(R), 1: rng.randrange(R), 2: rng.randrange(R)}

    @settings(max_examples=60, deadline=None)
    @given(st.lists(tree, min_size=1, max_size=3), st.booleans())
    def check(gates, fold):
        cg = zk.host.compile_gates(gates, 4, y_challenge=0 if fold else None)
        consts = list(cg.consts)
        for ci, slot in cg.challenge_slots.items():
            consts[slot] = challenges[ci]
        prog = [(cg.words[4 * i] & 0xFF, cg.words[4 * i] >> 8, cg.words[4 * i + 1], cg.words[4 * i + 2], cg.words[4 * i + 3])
                for i in range(len(cg.words) // 4)]
        written = set()
        for (op, dst, a, b, c) in prog:
            nsrc = 1 if op in (3, 4, 5, 7) else (3 if op == 6 else 2)
            for sref in (a, b, c)[:nsrc]:
                if sref >> 28 == 1:
                    assert (sref & 0x0FFFFFFF) in written
            assert dst < zk.host.GATE_MAX_REGS
            written.add(dst)
        got = pyref.gate_eval(prog, consts, cg.rotations, cols, k, k)
        for row in range(n):
            vals = [_eval_expr(g, cols, row, n, challenges) for g in gates]
            if fold:
                want = 0
                for v in vals:
                    want = (want * challenges[0] + v) % R
            else:
                want = sum(vals) % R
            assert got[row] == want

    check()


def test_compile_gates_register_allocation(zk, pyref):
    """Operands are evaluated hungrier-first, so a long one-sided chain (the shape of a y-folded gate list or a Horner
    form) needs two registers whatever its depth, and a complete binary tree of depth d needs d."""
    e = ("query", 0, 0)
    for _ in range(200):
        e = ("product", ("sum", ("query", 0, 0), ("const", 3)), ("sum", ("const", 5), e))
    cg = zk.host.compile_gates([e], 1)
    assert max(w >> 8 for w in cg.words[0::4]) <= 3
    col = [7, 11]
    prog = [(cg.words[4 * i] & 0xFF, cg.words[4 * i] >> 8, cg.words[4 * i + 1], cg.words[4 * i + 2], cg.words[4 * i + 3])
            for i in range(len(cg.words) // 4)]
    got = pyref.gate_eval(prog, cg.consts, cg.rotations, [col], 1, 1)
    assert got == [_eval_expr(e, [col], r, 2, {}) for r in range(2)]

    def bushy(d):
        return ("query", 0, 0) if d == 0 else ("product", bushy(d - 1), bushy(d - 1))
    # a complete binary tree of depth d needs d registers; overflowing the 48-register file would take 2^48 leaves
    cg = zk.host.compile_gates([bushy(10)], 1)
    assert max(w >> 8 for w in cg.words[0::4]) <= 10


def test_h2mo_open_restatement_is_accepted_by_the_pinned_verifier_pipeline(pyref):
    """The checker's prover (h2mo_open) against the checker's verifier scalar pipeline (h2mo_q_eval_sets / f_eval / v, pinned to
    Halo2MultiOpenMSM.hs:26-42 by test_oracle_kats): with the SRS secret known, commit(p) = p(s) G and the pairing check
    e(pi, [s]G2) == e(right, G2) is s * pi == right in G1."""
    import random
    P = pyref
    R = P.R_MOD
    rnd = random.Random(21)
    n, s = 32, rnd.randrange(R)
    polys = [[rnd.randrange(R) for _ in range(n)] for _ in range(5)]
    w = P.omega(5)
    x = rnd.randrange(R)
    queries = [(0, x), (1, x), (1, x * w % R), (2, x), (0, x * w % R), (3, x * w % R), (3, x), (3, x * P.fr_inv(w) % R), (4, x)]
    commit = lambda c: P.g1_mul(P.G1_GEN, P.poly_eval(c, s))
    comms = [commit(p) for p in polys]
    evals = [P.poly_eval(polys[i], pt) for i, pt in queries]
    t = P.Transcript()
    for c in comms:
        t.common_point(c)
    proof = P.h2mo_open(polys, queries, t, commit)
    psets, members, ev = P.h2mo_sets([(i, pt, e) for (i, pt), e in zip(queries, evals)])
    assert len(psets) == 3 and len(proof) == 48 + 32 * 3 + 48
    # verifier
    t = P.Transcript(proof)
    for c in comms:
        t.common_point(c)
    x1, x2 = t.squeeze(), t.squeeze()
    f_comm = t.read_point()
    x3 = t.squeeze()
    pq = [t.read_scalar() for _ in psets]
    x4 = t.squeeze()
    pi = t.read_point()
    cmap = [(comms[p], si, psets[si], ev[p]) for si, mem in enumerate(members) for p in mem]
    qes = P.h2mo_q_eval_sets(cmap, len(psets), x1)
    f_eval = P.h2mo_f_eval(psets, qes, x2, x3, pq)
    v = P.h2mo_v(f_eval, x4, pq)
    right = P.INF
    for si, mem in enumerate(members):
        xp = pow(x4, si, R)
        for p in mem:
            right = P.g1_add(right, P.g1_mul(comms[p], xp))
            xp = xp * x1 % R
    right = P.g1_add(right, P.g1_mul(f_comm, pow(x4, len(psets), R)))
    right = P.g1_add(right, P.g1_mul(P.g1_neg(P.G1_GEN), v))
    right = P.g1_add(right, P.g1_mul(pi, x3))
    assert P.g1_mul(pi, s) == right
    # and a wrong claimed evaluation is rejected
    ev2 = {k: list(vv) for k, vv in ev.items()}
    ev2[1][0] = (ev2[1][0] + 1) % R
    cmap2 = [(comms[p], si, psets[si], ev2[p]) for si, mem in enumerate(members) for p in mem]
    f2 = P.h2mo_f_eval(psets, P.h2mo_q_eval_sets(cmap2, len(psets), x1), x2, x3, pq)
    assert f2 != f_eval
