"""Times a batch of columns (the prover's commitment phase) through the resident entry point, for the setting of
B200ZK_BATCH_PIPE_MIN_POINTS in the environment (0 = the two-slot priority-stream pipeline off).  One JSON line per shape."""
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
    st = torch.cuda.current_stream().cuda_stream
    for k, ncol in ((14, 18), (17, 18), (19, 18), (20, 4)):
        n = 1 << k
        d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
        zk.capi.check(lib.b200zk_g1_synth_bases_dev(0xB200, 0, n, d_b.data_ptr(), st))
        torch.cuda.synchronize()
        h = C.c_uint64(0)
        zk.capi.check(lib.b200zk_bases_register_dev(d_b.data_ptr(), n, zk.FMT_MONT, 96, C.byref(h)))
        cols = [circuit_bench.prover_like(np, 100 + i, n) for i in range(ncol - 6)] + [bench.synth_scalars_np(200 + i, 0, n) for i in range(6)]
        d_sc = torch.from_numpy(np.concatenate(cols).view(np.uint8).reshape(-1)).cuda()
        d_out = torch.zeros(96 * ncol, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, ncol, 0, 0, d_out.data_ptr(), st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr(), n, ncol, 0, 0, d_out.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        # one uniform single MSM of the same size
        d_one = torch.zeros(96, dtype=torch.uint8, device="cuda")
        off = 32 * n * (ncol - 1)
        for _ in range(3):
            zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr() + off, n, 1, 0, 0, d_one.data_ptr(), st))
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(reps):
            zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_sc.data_ptr() + off, n, 1, 0, 0, d_one.data_ptr(), st))
        f1.record()
        torch.cuda.synchronize()
        prof = zk.capi.get_profile() if False else {}
        same = bytes(d_out[96 * (ncol - 1):].cpu().numpy()) == bytes(d_one.cpu().numpy())
        print(json.dumps({"pipe_min": os.environ.get("B200ZK_BATCH_PIPE_MIN_POINTS", "default"), "k": k, "columns": ncol,
                          "batch_ms": e0.elapsed_time(e1) / reps, "single_uniform_ms": f0.elapsed_time(f1) / reps,
                          "last_column_equals_single": same}), flush=True)
        zk.capi.check(lib.b200zk_bases_release(h.value))
        del d_b, d_sc


if __name__ == "__main__":
    main()
