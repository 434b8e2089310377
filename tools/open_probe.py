"""Per-repetition times of b200zk_h2mo_open_dev over resident random polynomials at several k (diagnostics).
usage: [B200ZK_H2MO_TIMING=1] python tools/open_probe.py [--ks 14,17,19] [--polys 15] [--reps 5]"""
import argparse
import ctypes as C
import importlib
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ks", default="14,17,19")
    ap.add_argument("--polys", type=int, default=15)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zk.init(0)
    lib = zk.lib()
    H = zk.host
    st = torch.cuda.current_stream().cuda_stream
    for k in [int(x) for x in args.ks.split(",")]:
        n = 1 << k
        rng = random.Random(k)
        d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
        zk.capi.check(lib.b200zk_g1_synth_bases_dev(0xB200, 0, n, d_b.data_ptr(), st))
        torch.cuda.synchronize()
        h = C.c_uint64(0)
        zk.capi.check(lib.b200zk_bases_register_dev(d_b.data_ptr(), n, zk.FMT_MONT, 96, C.byref(h)))
        d_cols = torch.from_numpy(bench.synth_scalars_np(300 + k, 0, n * args.polys).view(np.uint8).reshape(-1)).cuda()
        zk.capi.check(lib.b200zk_fr_convert_dev(d_cols.data_ptr(), d_cols.data_ptr(), n * args.polys, 1, st))
        torch.cuda.synchronize()
        w_k = pow(H.ROOT_OF_UNITY, 1 << (32 - k), H.R_MOD)
        x = rng.randrange(H.R_MOD)
        rots = [x, x * w_k % H.R_MOD, x * pow(w_k, H.R_MOD - 2, H.R_MOD) % H.R_MOD]
        queries = [(i, rots[0]) for i in range(args.polys)] + [(i, rots[1]) for i in range(0, args.polys, 3)] + \
                  [(i, rots[2]) for i in range(1, args.polys, 6)]
        ptrs = (C.c_void_p * args.polys)(*[d_cols.data_ptr() + 32 * n * i for i in range(args.polys)])
        qp = (C.c_uint32 * len(queries))(*[q[0] for q in queries])
        pts = b"".join(H.fr_bytes(q[1]) for q in queries)
        out = C.create_string_buffer(96 + 32 * len(queries))
        ln = C.c_size_t(0)
        times = []
        for _ in range(args.reps + 1):
            tr = H.Transcript()
            tr.common_scalar(k)
            t0 = time.perf_counter()
            zk.capi.check(lib.b200zk_h2mo_open_dev(h.value, tr.handle, C.addressof(ptrs), args.polys, n, C.addressof(qp), zk.capi.addr(pts),
                                                   len(queries), zk.capi.addr(out), len(out), C.byref(ln)))
            times.append((time.perf_counter() - t0) * 1e3)
            tr.free()
        print(json.dumps({"k": k, "polys": args.polys, "queries": len(queries), "first_ms": times[0], "rep_ms": times[1:]}), flush=True)
        zk.capi.check(lib.b200zk_bases_release(h.value))
        del d_b, d_cols


if __name__ == "__main__":
    main()
