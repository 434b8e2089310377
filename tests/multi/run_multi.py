"""Checks of the single-process multi-GPU path behind the plain C ABI (b200zk_init_devices): run as a script in a fresh
process by tests/test_gpu_multi.py (the test session itself is bound to one GPU).  Every result is compared with the
CPU oracle (test infrastructure).  Prints one JSON object; exit code 0 iff every check passed."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import threading

# one long transform is spread over the GPUs from 2^23 up; lowered so that the small sizes below take that path too
os.environ.setdefault("B200ZK_NTT_SHARD_MIN_LOG", "13")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--big-log-n", type=int, default=20)
    args = ap.parse_args()
    from conftest import Oracle, _build_oracle
    orc = Oracle(_build_oracle())
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    capi = zk.capi
    L = zk.lib()
    capi.init_devices(None, args.gpus)
    assert capi.device_count() == args.gpus
    res = {"gpus": args.gpus, "checks": []}

    def ok(name, cond):
        res["checks"].append({"name": name, "ok": bool(cond)})
        if not cond:
            print("FAILED:", name, file=sys.stderr)

    def register(points, n, flags):
        h = C.c_uint64(0)
        capi.check(L.b200zk_bases_register(capi.addr(points), n, flags, 96, C.byref(h)))
        return h.value

    def msm(h, off, sc, n, batch=1):
        out = C.create_string_buffer(96 * batch)
        capi.check(L.b200zk_msm_g1_batch(h, off, capi.addr(sc), n, batch, 0, capi.addr(out)))
        return [out.raw[96 * i:96 * i + 96] for i in range(batch)]

    n = (1 << 14) + 37
    pts = orc.synth_bases(0xB200, 0, n)
    cols = [orc.synth_scalars(10 + j, 0, n) for j in range(7)]
    want = [orc.msm(pts, c, n) for c in cols]
    for name, flags in (("shard", capi.BASES_SHARD), ("replicate", capi.BASES_REPLICATE), ("shard_no_tables", capi.BASES_SHARD | capi.BASES_NO_WINDOW_TABLES),
                        ("auto", 0)):
        h = register(pts, n, flags)
        ok(name + ": single column", msm(h, 0, cols[0], n) == want[:1])
        ok(name + ": 7 columns contiguous", msm(h, 0, b"".join(cols), n, 7) == want)
        ptrs = (C.c_void_p * 7)(*[capi.addr(c) for c in cols])
        out = C.create_string_buffer(96 * 7)
        capi.check(L.b200zk_msm_g1_batch_ptrs(h, 0, C.addressof(ptrs), n, 7, 0, capi.addr(out)))
        ok(name + ": 7 columns by pointer", [out.raw[96 * i:96 * i + 96] for i in range(7)] == want)
        # a slice that starts inside the table and is shorter than it (spans some shards only)
        off, m = n // 3 + 5, n // 2
        ok(name + ": offset slice", msm(h, off, cols[1][:32 * m], m) == [orc.msm(pts[96 * off:96 * (off + m)], cols[1][:32 * m], m)])
        ok(name + ": one point", msm(h, n - 1, cols[2][:32], 1) == [orc.msm(pts[96 * (n - 1):], cols[2][:32], 1)])
        ok(name + ": empty", msm(h, 0, b"", 0) == [bytes(96)])
        back = C.create_string_buffer(96 * 300)
        capi.check(L.b200zk_bases_read(h, n // 2 - 150, 300, capi.addr(back)))
        ok(name + ": read back across shards", back.raw == pts[96 * (n // 2 - 150):96 * (n // 2 + 150)])
        capi.check(L.b200zk_bases_release(h))

    # fewer points than GPUs
    h = register(pts, 1, capi.BASES_SHARD)
    ok("one-point table", msm(h, 0, cols[0][:32], 1) == [orc.msm(pts[:96], cols[0][:32], 1)])
    capi.check(L.b200zk_bases_release(h))

    # concurrent callers (rayon-style): four host threads against one sharded and one replicated table
    hs = register(pts, n, capi.BASES_SHARD)
    hr = register(pts, n, capi.BASES_REPLICATE)
    got = [None] * 8

    def worker(i):
        h = hs if i % 2 else hr
        got[i] = msm(h, 0, cols[i % 7], n)[0]

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    ok("8 concurrent host threads", got == [want[i % 7] for i in range(8)])

    # "_dev" entry point on a GPU other than the first: scalars allocated there, replicated table
    capi.set_device(args.gpus - 1)
    buf = zk.host.DeviceBuffer(32 * n)
    buf.upload(cols[3])
    dout = zk.host.DeviceBuffer(96)
    capi.check(L.b200zk_msm_g1_dev(hr, 0, buf.ptr, n, 1, 0, None, dout.ptr, None))
    ok("_dev on the last GPU (replicated table)", dout.download(96) == want[3])
    rc = L.b200zk_msm_g1_dev(hs, 0, buf.ptr, n, 1, 0, None, dout.ptr, None)
    ok("_dev against a slice that is not resident fails loudly", rc == (-1 if args.gpus > 1 else 0))
    capi.set_device(0)
    capi.check(L.b200zk_bases_release(hs))
    capi.check(L.b200zk_bases_release(hr))

    # the verifier's ad-hoc sum (bases not resident), long enough to be split by point range over the GPUs
    na = 20000
    apts = orc.synth_bases(0xB600, 0, na)
    asc = orc.synth_scalars(31, 0, na)
    out = C.create_string_buffer(96)
    capi.check(L.b200zk_msm_g1_adhoc(capi.addr(apts), 0, capi.addr(asc), 0, na, capi.addr(out)))
    ok("ad-hoc sum of 20000 points over all GPUs", out.raw == orc.msm(apts, asc, na))
    bad = bytearray(apts)
    bad[96 * (na - 3)] ^= 1
    rc = L.b200zk_msm_g1_adhoc(capi.addr(bytes(bad)), 0, capi.addr(asc), 0, na, capi.addr(out))
    ok("ad-hoc sum rejects a bad point on the last GPU's slice", rc == -7)

    # NTT batch dealt out over the GPUs (contiguous and by pointer), against the oracle
    k = 12
    omega = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - k), zk.host.R_MOD).to_bytes(32, "little")
    polys = [orc.synth_scalars(50 + j, 0, 1 << k) for j in range(5)]
    wantn = [orc.ntt(p, k, omega) for p in polys]
    blob = bytearray(b"".join(polys))
    capi.check(L.b200zk_ntt_fr_batch(capi.addr(blob), 5, k, capi.addr(omega), 0, None))
    ok("ntt batch contiguous", bytes(blob) == b"".join(wantn))
    bufs = [bytearray(p) for p in polys]
    ptrs = (C.c_void_p * 5)(*[capi.addr(b) for b in bufs])
    capi.check(L.b200zk_ntt_fr_batch_ptrs(C.addressof(ptrs), 5, k, capi.addr(omega), 0, None))
    ok("ntt batch by pointer", [bytes(b) for b in bufs] == wantn)

    # ONE transform over all GPUs (four-step, exchange fused into the column pass as peer stores): every flag combination at
    # two sizes against the oracle, the BASELINE-adjacent size 2^23 forward, and the resident form with its block layouts
    if args.gpus > 1:
        R_MOD = zk.host.R_MOD
        for kk in (13, 16, 20):
            nn = 1 << kk
            w = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - kk), R_MOD)
            wi = pow(w, R_MOD - 2, R_MOD)
            data = orc.synth_scalars(70 + kk, 0, nn)
            # (byte strings passed by address are kept in named variables so that they outlive the call)
            wb, wib, g7, g7i = (x.to_bytes(32, "little") for x in (w, wi, 7, pow(7, R_MOD - 2, R_MOD)))
            buf = bytearray(data)
            capi.check(L.b200zk_ntt_fr(capi.addr(buf), kk, capi.addr(wb), 0, None))
            fwd = bytes(buf)
            ok("sharded ntt 2^%d forward" % kk, fwd == orc.ntt(data, kk, wb))
            capi.check(L.b200zk_ntt_fr(capi.addr(buf), kk, capi.addr(wib), capi.NTT_INVERSE_SCALE, None))
            ok("sharded ntt 2^%d inverse" % kk, bytes(buf) == data)
            buf = bytearray(data)
            capi.check(L.b200zk_ntt_fr(capi.addr(buf), kk, capi.addr(wb), capi.NTT_COSET_IN, capi.addr(g7)))
            cos = bytes(buf)
            ok("sharded ntt 2^%d coset in" % kk, cos == orc.ntt(data, kk, wb, 0, g7))
            capi.check(L.b200zk_ntt_fr(capi.addr(buf), kk, capi.addr(wib), capi.NTT_INVERSE_SCALE | capi.NTT_COSET_OUT, capi.addr(g7i)))
            ok("sharded ntt 2^%d coset out" % kk, bytes(buf) == data)
        kk = 16
        nn = 1 << kk
        lr, lc = C.c_uint32(0), C.c_uint32(0)
        capi.check(L.b200zk_ntt_sharded_layout(kk, C.byref(lr), C.byref(lc)))
        Rr, Cc = 1 << lr.value, 1 << lc.value
        Cl, Rl = Cc // args.gpus, Rr // args.gpus
        w = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - kk), R_MOD)
        data = orc.synth_scalars(99, 0, nn)
        want = orc.ntt(data, kk, w.to_bytes(32, "little"))
        import numpy as np
        x = np.frombuffer(data, dtype=np.uint8).reshape(Rr, Cc, 32)
        ins, outs = [], []
        for g in range(args.gpus):
            capi.set_device(g)
            blk = np.ascontiguousarray(x[:, g * Cl:(g + 1) * Cl, :]).tobytes()
            b_in = zk.host.DeviceBuffer(len(blk))
            b_in.upload(blk)
            ins.append(b_in)
            outs.append(zk.host.DeviceBuffer(len(blk)))
        capi.set_device(0)
        pin = (C.c_void_p * args.gpus)(*[b.ptr for b in ins])
        pout = (C.c_void_p * args.gpus)(*[b.ptr for b in outs])
        wb = w.to_bytes(32, "little")
        capi.check(L.b200zk_ntt_fr_sharded_dev(C.addressof(pin), C.addressof(pout), args.gpus, kk, capi.addr(wb), 0, None))
        y = np.zeros((Cc, Rr, 32), dtype=np.uint8)            # X[k1 + R k2] at [k2][k1]
        for g in range(args.gpus):
            capi.set_device(g)
            y[:, g * Rl:(g + 1) * Rl, :] = np.frombuffer(outs[g].download(), dtype=np.uint8).reshape(Cc, Rl, 32)
        capi.set_device(0)
        ok("sharded ntt resident form (block layouts)", y.tobytes() == want)
        kk = 23
        data = orc.synth_scalars(5, 0, 1 << kk)
        w = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - kk), R_MOD).to_bytes(32, "little")
        buf = bytearray(data)
        capi.check(L.b200zk_ntt_fr(capi.addr(buf), kk, capi.addr(w), 0, None))
        ok("sharded ntt 2^23 forward vs oracle", bytes(buf) == orc.ntt(data, kk, w))

    # a larger sharded MSM in full against the oracle, and one by the discrete-log identity
    nb = 1 << args.big_log_n
    import numpy as np
    import torch
    d_b = torch.empty(96 * nb, dtype=torch.uint8, device="cuda:0")
    capi.check(L.b200zk_g1_synth_bases_dev(0xB200, 0, nb, d_b.data_ptr(), None))
    torch.cuda.synchronize()
    hb = C.c_uint64(0)
    capi.check(L.b200zk_bases_register_dev(d_b.data_ptr(), nb, capi.FMT_MONT | capi.BASES_SHARD, 96, C.byref(hb)))
    del d_b
    sc = orc.synth_scalars(1, 0, nb)
    got_big = msm(hb.value, 0, sc, nb)[0]
    ok("2^%d sharded from a device-side table vs oracle" % args.big_log_n, got_big == orc.msm(orc.synth_bases(0xB200, 0, nb), sc, nb))
    capi.check(L.b200zk_bases_release(hb.value))

    capi.shutdown()
    # shutdown -> init on another GPU -> NTT (per-device kernel attributes and cached tables must not survive)
    capi.init(args.gpus - 1)
    blob = bytearray(polys[0])
    capi.check(L.b200zk_ntt_fr(capi.addr(blob), k, capi.addr(omega), 0, None))
    ok("shutdown, init(last GPU), ntt", bytes(blob) == wantn[0])
    g, gl = zk.host.srs_generate(5, 6)
    ok("shutdown, init(last GPU), srs", zk.host.g1_export(g, 2)[:96] == orc.g1_generator())
    capi.shutdown()
    res["ok"] = all(c["ok"] for c in res["checks"])
    print(json.dumps(res))
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
