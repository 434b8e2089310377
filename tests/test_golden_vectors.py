"""The committed oracle-generated vectors (tests/golden/oracle_vectors.json, made by tests/golden/make_oracle_vectors.py):
on the CPU both oracles must still reproduce them; on the GPU the CUDA path must reproduce them through the C ABI."""
import ctypes as C
import hashlib
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


@pytest.fixture(scope="module")
def vec():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def fr_vec(P, seed, n):
    return [P.splitmix64(seed * 1000003 + i) * P.splitmix64(seed + 17 * i + 5) % R for i in range(n)]


def to_bytes(v):
    return b"".join(x.to_bytes(32, "little") for x in v)


def test_oracles_reproduce_the_vectors(vec, oracle, pyref):
    for m in vec["msm"]:
        n = m["n"]
        r = oracle.msm(oracle.synth_bases(0xB200, 0, n), oracle.synth_scalars(m["scalar_seed"], 0, n), n)
        assert oracle.g1_compress(r).hex() == m["result_compressed"]
    for t in vec["ntt"]:
        v = fr_vec(pyref, t["seed"], 1 << t["log_n"])
        w = pyref.omega(t["log_n"])
        assert sha(oracle.ntt(to_bytes(v), t["log_n"], w.to_bytes(32, "little"))) == t["forward_sha256"]
        assert sha(to_bytes(pyref.coset_ntt(v, w, 7))) == t["coset7_forward_sha256"]
    kd = vec["kate_div"]
    q, e = pyref.kate_div(fr_vec(pyref, kd["seed"], kd["n"]), int(kd["z"], 16))
    assert hex(e) == kd["eval"] and sha(to_bytes(q)) == kd["quotient_sha256"]


@pytest.mark.gpu
def test_cuda_path_reproduces_the_vectors(vec, gpu, oracle, pyref):
    H = gpu.host
    lib, chk, addr = gpu.lib(), gpu.capi.check, gpu.capi.addr
    # MSM through the resident-table entry point
    for m in vec["msm"]:
        n = m["n"]
        bases = oracle.synth_bases(0xB200, 0, n)
        h = C.c_uint64(0)
        chk(lib.b200zk_bases_register(addr(bases), n, gpu.FMT_CANONICAL, 96, C.byref(h)))
        out = C.create_string_buffer(96)
        sc = oracle.synth_scalars(m["scalar_seed"], 0, n)
        chk(lib.b200zk_msm_g1(h.value, 0, addr(sc), n, gpu.FMT_CANONICAL, addr(out)))
        chk(lib.b200zk_bases_release(h.value))
        assert H.g1_compress(out.raw).hex() == m["result_compressed"], n
    # NTT, plain and coset
    for t in vec["ntt"]:
        log_n = t["log_n"]
        v = to_bytes(fr_vec(pyref, t["seed"], 1 << log_n))
        wb = pyref.omega(log_n).to_bytes(32, "little")
        buf = bytearray(v)
        chk(lib.b200zk_ntt_fr(addr(buf), log_n, addr(wb), 0, None))
        assert sha(bytes(buf)) == t["forward_sha256"], log_n
        assert hex(int.from_bytes(buf[:32], "little")) == t["first"] and hex(int.from_bytes(buf[-32:], "little")) == t["last"]
        buf = bytearray(v)
        g7 = (7).to_bytes(32, "little")
        chk(lib.b200zk_ntt_fr(addr(buf), log_n, addr(wb), gpu.NTT_COSET_IN, addr(g7)))
        assert sha(bytes(buf)) == t["coset7_forward_sha256"], log_n
    # polynomial side
    kd = vec["kate_div"]
    q, e = H.fr_kate_div(H.FrVec.from_ints(fr_vec(pyref, kd["seed"], kd["n"])), int(kd["z"], 16))
    assert hex(e) == kd["eval"] and sha(q.to_canonical()) == kd["quotient_sha256"]
    rp = vec["running_product"]
    v = fr_vec(pyref, rp["seed"], rp["n"])
    assert sha(H.fr_running_product(H.FrVec.from_ints(v)).to_canonical()) == rp["exclusive_sha256"]
    assert hex(H.fr_running_product(H.FrVec.from_ints(v), inclusive=True).to_ints()[-1]) == rp["last_inclusive"]
    bi = vec["batch_invert"]
    vz = list(v)
    for i in bi["zeros_at"]:
        vz[i] = 0
    assert sha(H.fr_batch_invert(H.FrVec.from_ints(vz)).to_canonical()) == bi["sha256"]
    gpv = vec["gate_program"]
    cols = [H.FrVec.from_ints(fr_vec(pyref, s, 1 << gpv["ext_k"])) for s in gpv["column_seeds"]]
    consts = [fr_vec(pyref, gpv["const_seed"], 1)[0], 9]
    prog = H.GateProgram(gpv["words"], consts, gpv["rotations"], len(cols), gpv["k"], gpv["ext_k"],
                         pyref.vanishing_inverse_on_coset(7, gpv["k"], gpv["ext_k"]))
    out = prog.run(cols)
    assert sha(out.to_canonical()) == gpv["sha256"] and hex(out.to_ints()[0]) == gpv["row0"]
    prog.release()
    # SRS
    sv = vec["srs"]
    g, gl = H.srs_generate(int(sv["secret"], 16), sv["k"])
    n = 1 << sv["k"]
    wg, wl = H.g1_export(g, n), H.g1_export(gl, n)
    assert [H.g1_compress(wg[96 * i:96 * i + 96]).hex() for i in range(n)] == sv["g"]
    assert [H.g1_compress(wl[96 * i:96 * i + 96]).hex() for i in range(n)] == sv["g_lagrange"]
