"""Where a prover's commitment batch spends its time: sort / accumulate / tail (CUDA events inside the library) for batches of
prover-like (S) and uniform (U) columns at k = 17 / 19, resident.  One JSON line per batch shape (diagnostics)."""
import ctypes as C
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import circuit_bench  # noqa: E402


def main():
    import numpy as np
    import torch
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zk.init(0)
    lib = zk.lib()
    zk.capi.set_profiling(True)
    st = torch.cuda.current_stream().cuda_stream
    ks = [int(x) for x in os.environ.get("PROBE_KS", "17,19").split(",")]
    cbits = int(os.environ.get("PROBE_C", "0"))              # force the table's window bits (0 = the library's choice)
    for k in ks:
        n = 1 << k
        lib.b200zk_set_msm_tuning(cbits, 0)
        d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
        zk.capi.check(lib.b200zk_g1_synth_bases_dev(0xB200, 0, n, d_b.data_ptr(), st))
        torch.cuda.synchronize()
        h = C.c_uint64(0)
        zk.capi.check(lib.b200zk_bases_register_dev(d_b.data_ptr(), n, zk.FMT_MONT, 96, C.byref(h)))
        S = [circuit_bench.prover_like(np, 100 + i, n) for i in range(15)]
        U = [bench.synth_scalars_np(200 + i, 0, n) for i in range(3)]
        only = os.environ.get("PROBE_ONLY")                   # e.g. "15S+3U": one shape (for an ncu launch list)
        for label, cols in (("15S+3U", S + U), ("15S", S), ("3U", U), ("1S", S[:1]), ("1U", U[:1]), ("5S", S[:5])):
            if only and label != only:
                continue
            ncol = len(cols)
            d_sc = torch.from_numpy(np.concatenate(cols).view(np.uint8).reshape(-1)).cuda()
            d_out = torch.zeros(96 * ncol, dtype=torch.uint8, device="cuda")
            for _ in range(2):
                zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, ncol, 0, 0, d_out.data_ptr(), st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, ncol, 0, 0, d_out.data_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            prof = zk.capi.get_profile()
            print(json.dumps({"k": k, "batch": label, "ms": e0.elapsed_time(e1) / 5, **{x: prof.get(x) for x in ("sort", "accumulate", "tail", "window_bits", "windows")}}), flush=True)
            del d_sc
        zk.capi.check(lib.b200zk_bases_release(h.value))
        del d_b


if __name__ == "__main__":
    main()
