"""Bit-exact parity of the CUDA G1 MSM (through the C ABI) with the oracle: ragged sizes, the
edge cases the domain has (zero / maximal scalars, over-range scalars, repeated and opposite
points, identity points), skewed prover-like scalars, every wire format, batches, the verifier's
ad-hoc MSM, and -- at BASELINE sizes -- the discrete-log property MSM(s, a_i G) = (sum s_i a_i) G."""
import ctypes as C
import random

import pytest

pytestmark = pytest.mark.gpu

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def fr(x):
    return (x % R).to_bytes(32, "little")


NO_TABLES = 0x100   # B200ZK_BASES_NO_WINDOW_TABLES


def register(gpu, pts: bytes, n, fmt=0, stride=0):
    h = C.c_uint64(0)
    gpu.capi.check(gpu.lib().b200zk_bases_register(gpu.capi.addr(pts), n, fmt, stride, C.byref(h)))
    return h.value


def msm(gpu, h, sc: bytes, n, fmt=0, offset=0, batch=1):
    out = C.create_string_buffer(96 * batch)
    gpu.capi.check(gpu.lib().b200zk_msm_g1_batch(h, offset, gpu.capi.addr(sc), n, batch, fmt, gpu.capi.addr(out)))
    return out.raw


@pytest.fixture(scope="module")
def table(gpu, oracle):
    n = (1 << 14) + 3
    pts = oracle.synth_bases(0xB200, 0, n)
    h = register(gpu, pts, n)
    yield pts, h, n
    gpu.capi.check(gpu.lib().b200zk_bases_release(h))


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 32, 33, 100, 257, 1000, 4096, (1 << 14) + 3])
def test_uniform_scalars_vs_oracle(gpu, oracle, table, n):
    pts, h, _ = table
    sc = oracle.synth_scalars(1, 1000 * n, n)
    assert msm(gpu, h, sc, n) == oracle.msm(pts, sc, n)


def test_bases_roundtrip_and_offset(gpu, oracle, table):
    pts, h, n = table
    out = C.create_string_buffer(96 * 10)
    gpu.capi.check(gpu.lib().b200zk_bases_read(h, 5, 10, gpu.capi.addr(out)))
    assert out.raw == pts[96 * 5:96 * 15]
    sc = oracle.synth_scalars(4, 0, 300)
    assert msm(gpu, h, sc, 300, offset=1000) == oracle.msm(pts[96 * 1000:], sc, 300)
    assert gpu.lib().b200zk_msm_g1(h, n - 10, gpu.capi.addr(sc), 11, 0, gpu.capi.addr(out)) == -1
    assert gpu.lib().b200zk_msm_g1(h + 12345, 0, gpu.capi.addr(sc), 1, 0, gpu.capi.addr(out)) == -4


def test_empty_and_zero(gpu, oracle, table):
    pts, h, _ = table
    assert msm(gpu, h, b"", 0) == bytes(96)
    assert msm(gpu, h, bytes(32 * 100), 100) == bytes(96)


def test_edge_scalars(gpu, oracle, pyref, table):
    pts, h, _ = table
    n = 600
    rnd = random.Random(7)
    vals = [0, 1, 2, R - 1, R - 2, R, R + 1, (1 << 256) - 1, 1 << 255, (1 << 255) - 1, 1 << 128, (1 << 16) - 1, 1 << 15, (1 << 15) + 1]
    vals += [rnd.choice(vals) for _ in range(100)]
    vals += [rnd.randrange(R) for _ in range(n - len(vals))]
    raw = b"".join(v.to_bytes(32, "little") for v in vals)       # includes over-range encodings
    assert msm(gpu, h, raw, n) == oracle.msm(pts, raw, n)
    allmax = fr(R - 1) * n                                       # distribution "A": every scalar r-1
    assert msm(gpu, h, allmax, n) == oracle.msm(pts, allmax, n)
    same = fr(0x1234567) * n                                     # one bucket per window gets everything
    assert msm(gpu, h, same, n) == oracle.msm(pts, same, n)


def test_edge_points(gpu, oracle, pyref):
    """repeated points, P and -P, identity points (0,0), all in one table"""
    rnd = random.Random(8)
    G = pyref.G1_GEN
    base = [pyref.g1_mul(G, rnd.randrange(1, R)) for _ in range(40)]
    pts = []
    for i in range(400):
        q = base[i % 40]
        if i % 7 == 0:
            q = pyref.g1_neg(q)
        if i % 13 == 0:
            q = pyref.INF
        pts.append(q)
    sc = [rnd.randrange(R) for _ in range(400)]
    for i in range(0, 400, 40):
        sc[i] = 5                     # same point, same scalar -> exact doubling inside a bucket
    sc[1], sc[41] = 9, 9
    pb = b"".join(pyref.g1_to_wire(p) for p in pts)
    sb = b"".join(fr(s) for s in sc)
    h = register(gpu, pb, 400)
    got = msm(gpu, h, sb, 400)
    assert got == oracle.msm(pb, sb, 400) == pyref.g1_to_wire(pyref.g1_msm_naive(pts, sc))
    # P + (-P) with equal scalars cancels to the identity
    two = pyref.g1_to_wire(base[0]) + pyref.g1_to_wire(pyref.g1_neg(base[0]))
    h2 = register(gpu, two, 2)
    assert msm(gpu, h2, fr(77) * 2, 2) == bytes(96)
    for hh in (h, h2):
        gpu.capi.check(gpu.lib().b200zk_bases_release(hh))


@pytest.mark.parametrize("variant", [3, 4, 5, 6])
def test_edge_cases_every_accumulate_path(gpu, oracle, pyref, variant):
    """repeated / opposite / identity points and all-equal scalars through the XYZZ path (3) and the
    batched-affine paths (4, 5), with buckets long enough for several affine rounds"""
    rnd = random.Random(40 + variant)
    G = pyref.G1_GEN
    base = [pyref.g1_mul(G, rnd.randrange(1, R)) for _ in range(12)]
    pts = []
    for i in range(2000):
        q = base[i % 12]
        if i % 5 == 0:
            q = pyref.g1_neg(q)
        if i % 17 == 0:
            q = pyref.INF
        pts.append(q)
    pb = b"".join(pyref.g1_to_wire(p) for p in pts)
    try:
        gpu.capi.check(gpu.lib().b200zk_set_msm_tuning(8 | ((variant + 1) << 8), 512))
        h = register(gpu, pb, 2000)
        for sc in (fr(3) * 2000, fr(R - 1) * 2000, oracle.synth_scalars(50 + variant, 0, 2000),
                   b"".join(fr(rnd.randrange(4)) for _ in range(2000))):
            assert msm(gpu, h, sc, 2000) == oracle.msm(pb, sc, 2000)
        gpu.capi.check(gpu.lib().b200zk_bases_release(h))
    finally:
        gpu.lib().b200zk_set_msm_tuning(0, 0)


def test_prover_like_skew(gpu, oracle, table):
    """distribution "S": 70% in {0,1}, 20% < 2^16, 10% uniform -- the heavy-bucket split path"""
    pts, h, n = table
    rnd = random.Random(9)
    uni = oracle.synth_scalars(6, 0, n)
    vals = []
    for i in range(n):
        u = rnd.random()
        if u < 0.7:
            vals.append(fr(rnd.randrange(2)))
        elif u < 0.9:
            vals.append(fr(rnd.randrange(1 << 16)))
        else:
            vals.append(uni[32 * i:32 * i + 32])
    sc = b"".join(vals)
    assert msm(gpu, h, sc, n) == oracle.msm(pts, sc, n)


@pytest.mark.parametrize("c,smax", [(4, 8), (7, 16), (10, 0), (13, 8), (16, 0), (18, 64)])
@pytest.mark.parametrize("tables", [False, True])
def test_all_window_sizes_and_task_splits(gpu, oracle, table, c, smax, tables):
    """every window width, with and without the precomputed window tables, forced task splits,
    and every accumulate-kernel code variant"""
    pts, _, _ = table
    n = 3000
    sc = oracle.synth_scalars(10 + c, 0, n)
    exp = oracle.msm(pts, sc, n)
    try:
        if tables and c < 8:
            pytest.skip("window tables need c >= 8")
        gpu.capi.check(gpu.lib().b200zk_set_msm_tuning(c, smax))
        h = register(gpu, pts, n, fmt=0 if tables else NO_TABLES)
        assert msm(gpu, h, sc, n) == exp
        assert msm(gpu, h, sc[32 * 100:], n - 500, offset=100) == oracle.msm(pts[96 * 100:], sc[32 * 100:], n - 500)
        for variant in range(7):
            gpu.capi.check(gpu.lib().b200zk_set_msm_tuning(c | ((variant + 1) << 8), smax))
            assert msm(gpu, h, sc, n) == exp, variant
        gpu.capi.check(gpu.lib().b200zk_bases_release(h))
    finally:
        gpu.lib().b200zk_set_msm_tuning(0, 0)


def test_montgomery_formats_and_stride(gpu, oracle, pyref):
    n = 200
    pts = oracle.synth_bases(0xB200, 50, n)
    sc = oracle.synth_scalars(12, 0, n)
    exp = oracle.msm(pts, sc, n)
    sc_mont = b"".join(fr(int.from_bytes(sc[32 * i:32 * i + 32], "little") << 256) for i in range(n))
    tomont = lambda b: ((int.from_bytes(b, "little") << 384) % P).to_bytes(48, "little")
    # blst_p1_affine layout inside a wider struct: x, y, then 8 bytes of padding (stride 104)
    wide = b"".join(tomont(pts[96 * i:96 * i + 48]) + tomont(pts[96 * i + 48:96 * i + 96]) + b"\xAA" * 8 for i in range(n))
    h = register(gpu, wide, n, fmt=gpu.FMT_MONT, stride=104)
    assert msm(gpu, h, sc, n) == exp
    assert msm(gpu, h, sc_mont, n, fmt=gpu.FMT_MONT) == exp
    gpu.capi.check(gpu.lib().b200zk_bases_release(h))


def test_bad_points_rejected(gpu, pyref):
    G = pyref.g1_to_wire(pyref.G1_GEN)
    off_curve = G[:48] + (int.from_bytes(G[48:], "little") + 1).to_bytes(48, "little")
    h = C.c_uint64(0)
    assert gpu.lib().b200zk_bases_register(gpu.capi.addr(G + off_curve), 2, 0, 0, C.byref(h)) == -7
    non_canon = (int.from_bytes(G[:48], "little") + P).to_bytes(48, "little") + G[48:]
    assert gpu.lib().b200zk_bases_register(gpu.capi.addr(non_canon), 1, 0, 0, C.byref(h)) == -7
    assert gpu.lib().b200zk_bases_register(gpu.capi.addr(G), 1, 0, 50, C.byref(h)) == -1


def test_batch_matches_singles(gpu, oracle, table):
    pts, h, _ = table
    n, batch = 1500, 6
    sc = oracle.synth_scalars(13, 0, n * batch)
    got = msm(gpu, h, sc, n, batch=batch)
    for b in range(batch):
        assert got[96 * b:96 * b + 96] == oracle.msm(pts, sc[32 * n * b:32 * n * (b + 1)], n), b


def test_commitment_scheme_mirror(gpu, oracle):
    """ParamsKZG / KZGCommitmentScheme.commit / commit_lagrange with the reference's shapes
    (simple_mul: k = 5, ten commitments per proof)."""
    k = 5
    n = 1 << k
    g = oracle.synth_bases(0xB200, 0, n)
    gl = oracle.synth_bases(0xB201, 0, n)
    params = gpu.host.ParamsKZG(k, g, gl)
    polys = [oracle.synth_scalars(20 + i, 0, n) for i in range(10)]
    K = gpu.host.KZGCommitmentScheme
    assert K.commit(params, polys[0]) == oracle.msm(g, polys[0], n)
    assert K.commit_lagrange(params, polys[1]) == oracle.msm(gl, polys[1], n)
    assert K.commit_batch(params, polys, lagrange=True) == [oracle.msm(gl, p, n) for p in polys]
    with pytest.raises(AssertionError):
        K.commit(params, polys[0] + polys[1])
    params.release()


def test_verifier_final_msm_from_reference_fixture(gpu, oracle, pyref, kats):
    """Right-hand MSM of the multi-open verifier built from the reference's ProofData fixture
    (Halo2MultiOpenMSM.hs:60-98): sum_s x4^s sum_j x1^j C_{s,j} + x4^S f - v G + x3 pi."""
    h = kats["h2mo"]
    S = {k: int(v, 16) for k, v in h["scalars"].items()}
    PT = {k: pyref.g1_to_wire((int(v[0], 16), int(v[1], 16))) for k, v in h["points"].items()}
    v = int(h["expected_v"], 16)
    msm_ = gpu.host.DualMSM()
    for s in range(3):
        xp = 1
        for c in h["commitment_map"]:
            if c["set"] != s:
                continue
            msm_.append_right(pow(S["x4"], s, R) * xp % R, PT[c["commitment"]])
            xp = xp * S["x1"] % R
    f_comm = PT.get("f_commitment", PT["a1"])      # fixture has no f/pi point: any on-curve point exercises the sum
    msm_.append_right(pow(S["x4"], 3, R), f_comm)
    msm_.append_right(-v, pyref.g1_to_wire(pyref.G1_GEN))
    msm_.append_right(S["x3"], PT["a2"])
    msm_.append_left(1, PT["a2"])
    left, right = msm_.eval()
    pts = b"".join(p for _, p in msm_.right)
    sc = b"".join(fr(s) for s, _ in msm_.right)
    assert right == oracle.msm(pts, sc, len(msm_.right), naive=True)
    assert left == PT["a2"]


def test_batched_verifier_shape(gpu, oracle):
    """config 5 shape at reduced scale: ad-hoc MSM over 64 proofs x 26 points"""
    n = 64 * 26
    pts = oracle.synth_bases(0xB300, 0, n)
    sc = oracle.synth_scalars(14, 0, n)
    out = C.create_string_buffer(96)
    gpu.capi.check(gpu.lib().b200zk_msm_g1_adhoc(gpu.capi.addr(pts), 0, gpu.capi.addr(sc), 0, n, gpu.capi.addr(out)))
    assert out.raw == oracle.msm(pts, sc, n)


def test_synthetic_bases_and_dlog_property_2pow20(gpu, oracle, pyref):
    """Device-resident path at a BASELINE size: bases generated on the GPU (spot-checked against the
    oracle), scalars resident in HBM, result checked through the discrete-log identity."""
    import numpy as np
    import torch
    n = 1 << 20
    seed = 0xB200
    d_bases = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gpu.capi.check(gpu.lib().b200zk_g1_synth_bases_dev(seed, 0, n, d_bases.data_ptr(), st))
    torch.cuda.synchronize()
    h = C.c_uint64(0)
    gpu.capi.check(gpu.lib().b200zk_bases_register_dev(d_bases.data_ptr(), n, gpu.FMT_MONT, 96, C.byref(h)))
    out = C.create_string_buffer(96 * 4)
    for start in (0, 12345, n - 4):
        gpu.capi.check(gpu.lib().b200zk_bases_read(h.value, start, 4, gpu.capi.addr(out)))
        assert out.raw == oracle.synth_bases(seed, start, 4)
    sc = oracle.synth_scalars(1, 0, n)
    d_sc = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    gpu.capi.check(gpu.lib().b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, 1, 0, 0, d_out.data_ptr(), st))
    torch.cuda.synchronize()
    got = bytes(d_out.cpu().numpy())
    a = np.array([pyref.synth_base_dlog(seed, i) for i in range(n)], dtype=np.uint64)
    dot = C.create_string_buffer(32)
    oracle.L.orc_fr_dot_u64(sc, a.ctypes.data, n, dot)
    assert got == oracle.g1_mul(oracle.g1_generator(), dot.raw)
    # the same through the host-buffer entry point
    assert msm(gpu, h.value, sc, n) == got
    gpu.capi.check(gpu.lib().b200zk_bases_release(h.value))


def test_g1_sum_dev(gpu, oracle):
    """combine step of the point-range sharded MSM: sum of per-GPU partial results"""
    import torch
    n = 8
    pts = oracle.synth_bases(0xB400, 0, n)
    P_ = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    mont = b"".join(((int.from_bytes(pts[48 * i:48 * i + 48], "little") << 384) % P_).to_bytes(48, "little") for i in range(2 * n))
    d = torch.frombuffer(bytearray(mont), dtype=torch.uint8).cuda()
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gpu.capi.check(gpu.lib().b200zk_g1_sum_dev(d.data_ptr(), n, 0, d_out.data_ptr(), st))
    torch.cuda.synchronize()
    assert bytes(d_out.cpu().numpy()) == oracle.g1_sum(pts, n)


def test_sharded_partials_emulated_on_one_gpu(gpu, oracle):
    """The N > 1 data path on one device: two 'ranks' each hold a slice of the table, compute an
    un-normalised XYZZ partial (b200zk_msm_g1_partial_dev), the partials are concatenated as the
    all-gather would and folded by b200zk_g1_sum_partials_dev."""
    import importlib
    import torch
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    n, world = 5001, 2
    pts = oracle.synth_bases(0xB500, 0, n)
    sc = oracle.synth_scalars(21, 0, n)
    st = torch.cuda.current_stream().cuda_stream
    d_all = torch.zeros(192 * world, dtype=torch.uint8, device="cuda")
    handles = []
    for r in range(world):
        s, e = zdist.shard_range(n, r, world)
        h = register(gpu, pts[96 * s:96 * e], e - s)
        handles.append(h)
        d_sc = torch.frombuffer(bytearray(sc[32 * s:32 * e]), dtype=torch.uint8).cuda()
        gpu.capi.check(gpu.lib().b200zk_msm_g1_partial_dev(h, 0, d_sc.data_ptr(), e - s, 0, d_all[192 * r:].data_ptr(), st))
        torch.cuda.synchronize()
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    gpu.capi.check(gpu.lib().b200zk_g1_sum_partials_dev(d_all.data_ptr(), world, 0, d_out.data_ptr(), st))
    torch.cuda.synchronize()
    assert bytes(d_out.cpu().numpy()) == oracle.msm(pts, sc, n)
    # ShardedMSM with world = 1 goes through the same entry points
    m = zdist.ShardedMSM(handles[0], zdist.shard_range(n, 0, world)[1], 0, 1)
    s0, e0 = zdist.shard_range(n, 0, world)
    d_sc = torch.frombuffer(bytearray(sc[:32 * e0]), dtype=torch.uint8).cuda()
    out = m.run_device(d_sc)
    torch.cuda.synchronize()
    assert bytes(out.cpu().numpy()) == oracle.msm(pts[:96 * e0], sc[:32 * e0], e0)
    for h in handles:
        gpu.capi.check(gpu.lib().b200zk_bases_release(h))


@pytest.mark.parametrize("dist", ["uniform", "prover_like", "constant"])
def test_large_input_sort_paths(gpu, oracle, dist):
    """2^19 points: large enough for the two-pass counting sort (uniform) and for its device-side
    fall-back to the one-pass sort on skewed inputs (0/1 columns, constant scalars)."""
    import numpy as np
    n = 1 << 19
    pts = oracle.synth_bases(0xB700, 0, n)
    h = register(gpu, pts, n)
    uni = np.frombuffer(oracle.synth_scalars(61, 0, n), dtype=np.uint64).reshape(n, 4).copy()
    if dist == "prover_like":
        rng = np.random.default_rng(5)
        u = rng.random(n)
        small = u < 0.9
        uni[small, 1:] = 0
        uni[small, 0] = np.where(u[small] < 0.7, rng.integers(0, 2, small.sum(), dtype=np.uint64),
                                 rng.integers(0, 1 << 16, small.sum(), dtype=np.uint64))
    elif dist == "constant":
        uni[:] = np.frombuffer((R - 1).to_bytes(32, "little"), dtype=np.uint64)
    sc = uni.tobytes()
    assert msm(gpu, h, sc, n) == oracle.msm(pts, sc, n)
    gpu.capi.check(gpu.lib().b200zk_bases_release(h))


def test_calls_on_different_streams_do_not_race(gpu, oracle):
    """The workspaces are shared by all calls of the process: back-to-back device-side calls on two different CUDA
    streams, never synchronised in between, must still give the single-stream results (the library orders them)."""
    import ctypes as C
    import torch
    lib, chk = gpu.lib(), gpu.capi.check
    n = 1 << 14
    bases = oracle.synth_bases(0xB200, 0, n)
    h = C.c_uint64(0)
    chk(lib.b200zk_bases_register(gpu.capi.addr(bases), n, gpu.FMT_CANONICAL, 96, C.byref(h)))
    sets = [oracle.synth_scalars(50 + i, 0, n) for i in range(4)]
    want = [oracle.msm(bases, s, n) for s in sets]
    d_sc = [torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda() for s in sets]
    d_out = [torch.zeros(96, dtype=torch.uint8, device="cuda") for _ in sets]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(4):
            st = streams[i & 1]
            chk(lib.b200zk_msm_g1_dev(h.value, 0, d_sc[i].data_ptr(), n, 1, 0, 0, d_out[i].data_ptr(), st.cuda_stream))
        torch.cuda.synchronize()
        for i in range(4):
            assert bytes(d_out[i].cpu().numpy()) == want[i], (rep, i)
    chk(lib.b200zk_bases_release(h.value))


def test_chunked_host_buffer_path_small_sizes(gpu):
    """A host-buffer MSM of >= B200ZK_MSM_CHUNK_MIN points streams its scalars in two halves (second bucket set, merge
    kernel).  The default threshold is 2^22; a child process lowers it to 512 so that ragged small sizes, skewed scalars
    and tables with and without window rows go through that path against the oracle."""
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
lib, chk, addr = zk.lib(), zk.capi.check, zk.capi.addr
R = zk.host.R_MOD
for n, flags in ((512, 0), (777, 0), (4099, 0), (4099, 0x100), (20000, 0)):
    bases = orc.synth_bases(0xB200, 5, n)
    h = C.c_uint64(0)
    chk(lib.b200zk_bases_register(addr(bases), n, zk.FMT_CANONICAL | flags, 96, C.byref(h)))
    for sc in (orc.synth_scalars(70 + n, 0, n), (R - 1).to_bytes(32, "little") * n,
               b"".join((i %% 2).to_bytes(32, "little") for i in range(n))):
        out = C.create_string_buffer(96)
        chk(lib.b200zk_msm_g1(h.value, 0, addr(sc), n, zk.FMT_CANONICAL, addr(out)))
        assert out.raw == orc.msm(bases, sc, n), (n, flags)
    chk(lib.b200zk_bases_release(h.value))
print("chunked ok")
''' % (root, root)
    env = dict(os.environ, B200ZK_MSM_CHUNK_MIN="512")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "chunked ok" in r.stdout, r.stdout + r.stderr


def test_table_window_bits_of_prover_size_tables(gpu, oracle):
    """The window of a table is fixed at registration: 16 bits for 2^12 < n <= 2^18, 18 at k = 19 (measured on
    commitment batches); whatever it is, the result is the oracle's."""
    gpu.capi.set_profiling(True)
    try:
        for k, want in ((13, 16), (16, 16), (19, 18)):
            n = 1 << k
            bases = oracle.synth_bases(0xB200, 0, n)
            h = register(gpu, bases, n)
            sc = oracle.synth_scalars(60 + k, 0, n)
            got = msm(gpu, h, sc, n)
            prof = gpu.capi.get_profile()
            assert prof["window_bits"] == want, (k, prof)
            assert got == oracle.msm(bases, sc, n), k
            gpu.capi.check(gpu.lib().b200zk_bases_release(h))
    finally:
        gpu.capi.set_profiling(False)


@pytest.mark.parametrize("env", [{"B200ZK_RED_TP": "3", "B200ZK_RED_TP_MIN": "1"}, {"B200ZK_RED_TP": "4", "B200ZK_RED_TP_MIN": "1"},
                                 {"B200ZK_COOP8_MAX": "100000"}, {"B200ZK_COOP8_MAX": "1"}])
def test_bucket_tree_kernel_choices_agree(gpu, env):
    """The bucket tree picks its kernels by group count (serial / throughput build / warp per group / CTA per group); a child
    process forces each choice at sizes where the default would not take it -- same commitments, with and without window
    tables, single columns and a batch, skewed scalars included."""
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
lib, chk, addr = zk.lib(), zk.capi.check, zk.capi.addr
R = zk.host.R_MOD
for n, flags in ((300, 0), (5000, 0), (5000, 0x100), (70001, 0)):
    bases = orc.synth_bases(0xB200, 9, n)
    h = C.c_uint64(0)
    chk(lib.b200zk_bases_register(addr(bases), n, zk.FMT_CANONICAL | flags, 96, C.byref(h)))
    cols = [orc.synth_scalars(90 + n, 0, n), (R - 1).to_bytes(32, "little") * n,
            b"".join((i %% 3).to_bytes(32, "little") for i in range(n))]
    out = C.create_string_buffer(96 * len(cols))
    chk(lib.b200zk_msm_g1_batch(h.value, 0, addr(b"".join(cols)), n, len(cols), zk.FMT_CANONICAL, addr(out)))
    for j, sc in enumerate(cols):
        want = orc.msm(bases, sc, n)
        assert out.raw[96 * j:96 * j + 96] == want, (n, flags, j, "batch")
        one = C.create_string_buffer(96)
        chk(lib.b200zk_msm_g1(h.value, 0, addr(sc), n, zk.FMT_CANONICAL, addr(one)))
        assert one.raw == want, (n, flags, j, "single")
    chk(lib.b200zk_bases_release(h.value))
print("tree ok")
''' % (root, root)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "tree ok" in r.stdout, r.stdout + r.stderr
