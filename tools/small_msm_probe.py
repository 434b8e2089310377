"""Latency probe for the small MSMs of the example circuits and the verifier (simple_mul k=5, lookup_table k=11,
atms k=14, the ad-hoc batched-verify sum): host-buffer e2e time, device time and phase split per (log_n, batch).
usage: python tools/small_msm_probe.py [--adhoc-only]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--adhoc-only", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import numpy as np
    import torch
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zk.init(0)
    lib, chk = zk.lib(), zk.capi.check
    zk.capi.set_profiling(True)
    st = torch.cuda.current_stream().cuda_stream
    if not args.adhoc_only:
        for log_n, batch in ((5, 1), (5, 10), (8, 10), (11, 1), (11, 20), (14, 1), (14, 18), (16, 1), (17, 18)):
            n = 1 << log_n
            d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
            chk(lib.b200zk_g1_synth_bases_dev(bench.BASE_SEED, 0, n, d_b.data_ptr(), st))
            torch.cuda.synchronize()
            h = C.c_uint64(0)
            chk(lib.b200zk_bases_register_dev(d_b.data_ptr(), n, zk.FMT_MONT, 96, C.byref(h)))
            h_sc = torch.from_numpy(bench.synth_scalars_np(1, 0, n * batch).view(np.uint8).reshape(-1)).pin_memory()
            d_sc = h_sc.cuda()
            d_out = torch.zeros(96 * batch, dtype=torch.uint8, device="cuda")
            out = torch.zeros(96 * batch, dtype=torch.uint8).pin_memory()
            for _ in range(3):
                chk(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, batch, 0, 0, d_out.data_ptr(), st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = zk.launch_count()
            e0.record()
            for _ in range(args.reps):
                chk(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, batch, 0, 0, d_out.data_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            dev_ms = e0.elapsed_time(e1) / args.reps
            launches = (zk.launch_count() - l0) // args.reps
            prof = zk.capi.get_profile()
            chk(lib.b200zk_msm_g1_batch(h.value, 0, h_sc.data_ptr(), n, batch, 0, out.data_ptr()))
            t0 = time.perf_counter()
            for _ in range(args.reps):
                chk(lib.b200zk_msm_g1_batch(h.value, 0, h_sc.data_ptr(), n, batch, 0, out.data_ptr()))
            e2e = (time.perf_counter() - t0) / args.reps * 1e3
            print(json.dumps({"op": "msm_resident", "log_n": log_n, "batch": batch, "dev_ms": dev_ms, "e2e_ms": e2e, "launches": launches,
                              "phases_ms": {x: prof.get(x) for x in ("sort", "accumulate", "tail")}, "window_bits": prof.get("window_bits"),
                              "windows": prof.get("windows")}), flush=True)
            chk(lib.b200zk_bases_release(h.value))
    for n in (1024, 27648):
        d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
        chk(lib.b200zk_g1_synth_bases_dev(bench.BASE_SEED, 0, n, d_b.data_ptr(), st))
        torch.cuda.synchronize()
        h_pts = d_b.cpu().pin_memory()
        h_sc = torch.from_numpy(bench.synth_scalars_np(9, 0, n).view(np.uint8).reshape(-1)).pin_memory()
        out = C.create_string_buffer(96)
        for _ in range(3):
            chk(lib.b200zk_msm_g1_adhoc(h_pts.data_ptr(), zk.FMT_MONT, h_sc.data_ptr(), 0, n, zk.capi.addr(out)))
        l0 = zk.launch_count()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            chk(lib.b200zk_msm_g1_adhoc(h_pts.data_ptr(), zk.FMT_MONT, h_sc.data_ptr(), 0, n, zk.capi.addr(out)))
        e2e = (time.perf_counter() - t0) / args.reps * 1e3
        prof = zk.capi.get_profile()
        print(json.dumps({"op": "msm_adhoc", "points": n, "e2e_ms": e2e, "launches": (zk.launch_count() - l0) // args.reps,
                          "phases_ms": {x: prof.get(x) for x in ("sort", "accumulate", "tail")}, "window_bits": prof.get("window_bits"),
                          "windows": prof.get("windows")}), flush=True)


if __name__ == "__main__":
    main()
