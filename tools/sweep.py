"""Reporting grid of SURVEY.md section 8d on one GPU: G1 MSM at 2^16..2^24 for the three scalar
distributions (U uniform, S prover-like, A all r-1) and the Fr NTT at 2^16..2^24 (forward, inverse,
coset forward), device-resident (CUDA events) and end to end (host buffers).  Writes one JSON line
per cell.  usage: python tools/sweep.py [--sizes 16,18,20,22,24] [--reps 5]"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from circuit_bench import prover_like  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16,18,20,22,24")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch

    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zk.init(0)
    lib = zk.lib()
    zk.capi.set_profiling(True)
    st = torch.cuda.current_stream().cuda_stream
    ops, ms_ = C.c_double(), C.c_double()
    zk.capi.check(lib.b200zk_microbench(7, 4000, C.byref(ops), C.byref(ms_)))
    imad_peak = ops.value
    R = bench.R_MOD
    for log_n in [int(x) for x in args.sizes.split(",")]:
        n = 1 << log_n
        d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
        zk.capi.check(lib.b200zk_g1_synth_bases_dev(bench.BASE_SEED, 0, n, d_b.data_ptr(), st))
        torch.cuda.synchronize()
        h = C.c_uint64(0)
        t0 = time.perf_counter()
        zk.capi.check(lib.b200zk_bases_register_dev(d_b.data_ptr(), n, zk.FMT_MONT, 96, C.byref(h)))
        reg_ms = (time.perf_counter() - t0) * 1e3
        del d_b
        torch.cuda.empty_cache()
        dists = {"U": bench.synth_scalars_np(1, 0, n), "S": prover_like(np, 7, n),
                 "A": np.tile(np.frombuffer((R - 1).to_bytes(32, "little"), dtype=np.uint64), (n, 1))}
        for name, k in dists.items():
            h_sc = torch.from_numpy(np.ascontiguousarray(k).view(np.uint8).reshape(-1)).pin_memory()
            d_sc = h_sc.cuda()
            d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
            out = C.create_string_buffer(96)
            for _ in range(3):
                zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, 1, 0, 0, d_out.data_ptr(), st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, 1, 0, 0, d_out.data_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            prof = zk.capi.get_profile()
            zk.capi.check(lib.b200zk_msm_g1(h.value, 0, h_sc.data_ptr(), n, 0, zk.capi.addr(out)))
            t0 = time.perf_counter()
            for _ in range(args.reps):
                zk.capi.check(lib.b200zk_msm_g1(h.value, 0, h_sc.data_ptr(), n, 0, zk.capi.addr(out)))
            e2e = (time.perf_counter() - t0) / args.reps * 1e3
            assert out.raw == bytes(d_out.cpu().numpy())
            print(json.dumps({"op": "msm_g1", "log_n": log_n, "dist": name, "ms": ms, "points_per_s": n / ms * 1e3,
                              "e2e_ms": e2e, "e2e_points_per_s": n / e2e * 1e3, "phases_ms": {x: prof.get(x) for x in ("sort", "accumulate", "tail")},
                              "window_bits": prof.get("window_bits"), "windows": prof.get("windows"),
                              "imad_frac_48k": (n * bench.LMAC_PER_POINT / (ms * 1e-3)) / imad_peak,
                              "table_build_ms": reg_ms, "result": zk.host.g1_compress(out.raw).hex()[:16]}), flush=True)
        zk.capi.check(lib.b200zk_bases_release(h.value))
        # ---- NTT
        w = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - log_n), R)
        wb, wib = w.to_bytes(32, "little"), pow(w, R - 2, R).to_bytes(32, "little")
        g = zk.host.ZETA.to_bytes(32, "little")
        h_d = torch.from_numpy(bench.synth_scalars_np(2, 0, n).view(np.uint8).reshape(-1)).pin_memory()
        d_d = h_d.cuda()
        for label, om, flags, shift in (("forward", wb, 0, None), ("inverse", wib, zk.NTT_INVERSE_SCALE, None),
                                        ("coset_forward", wb, zk.NTT_COSET_IN, g)):
            for _ in range(3):
                zk.capi.check(lib.b200zk_ntt_fr_dev(d_d.data_ptr(), 1, log_n, zk.capi.addr(om), flags, zk.capi.addr(shift), st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                zk.capi.check(lib.b200zk_ntt_fr_dev(d_d.data_ptr(), 1, log_n, zk.capi.addr(om), flags, zk.capi.addr(shift), st))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            zk.capi.check(lib.b200zk_ntt_fr(h_d.data_ptr(), log_n, zk.capi.addr(om), flags, zk.capi.addr(shift)))   # staging buffers grow once
            t0 = time.perf_counter()
            for _ in range(args.reps):
                zk.capi.check(lib.b200zk_ntt_fr(h_d.data_ptr(), log_n, zk.capi.addr(om), flags, zk.capi.addr(shift)))
            e2e = (time.perf_counter() - t0) / args.reps * 1e3
            passes = 1 if log_n <= 11 else (2 if log_n <= 22 else 3)
            lmacs = (n // 2) * log_n * bench.LMAC_PER_FR_MUL
            print(json.dumps({"op": "ntt_fr", "log_n": log_n, "variant": label, "ms": ms, "elements_per_s": n / ms * 1e3, "e2e_ms": e2e,
                              "hbm_gbs": passes * 64 * n / (ms * 1e-3) / 1e9, "imad_frac_136_per_product": (lmacs / (ms * 1e-3)) / imad_peak,
                              "imad_frac_issued": ((n // 2) * log_n * 113 / (ms * 1e-3)) / imad_peak,
                              "passes": passes}), flush=True)


if __name__ == "__main__":
    main()
