"""Times the device-resident Fr NTT at a few sizes for the kernel variant selected by B200ZK_NTT_VARIANT (one process per
variant) and checks the variant against the CPU checker at 2^16 (two passes) and 2^11 (one pass).  One JSON line per size."""
import ctypes as C
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import numpy as np
    import torch
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zk.init(0)
    lib = zk.lib()
    L = bench.load_oracle()
    variant = os.environ.get("B200ZK_NTT_VARIANT", "default")
    for k in (11, 16):
        n = 1 << k
        omega = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - k), bench.R_MOD).to_bytes(32, "little")
        ref = bench.synth_scalars_np(7, 0, n)
        got = torch.from_numpy(ref.copy().view(np.uint8).reshape(-1)).cuda()
        zk.capi.check(lib.b200zk_ntt_fr_dev(got.data_ptr(), 1, k, zk.capi.addr(omega), 0, 0, None))
        L.orc_ntt(ref.ctypes.data, k, omega, 0, None, None, 0)
        torch.cuda.synchronize()
        assert bytes(got.cpu().numpy()) == ref.tobytes(), "variant %s differs from the checker at 2^%d" % (variant, k)
    zk.capi.set_profiling(True)
    for k in (20, 22, 24):
        n = 1 << k
        omega = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - k), bench.R_MOD).to_bytes(32, "little")
        d = torch.from_numpy(bench.synth_scalars_np(2, 0, n).view(np.uint8).reshape(-1)).cuda()
        for _ in range(5):
            zk.capi.check(lib.b200zk_ntt_fr_dev(d.data_ptr(), 1, k, zk.capi.addr(omega), 0, 0, None))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            zk.capi.check(lib.b200zk_ntt_fr_dev(d.data_ptr(), 1, k, zk.capi.addr(omega), 0, 0, None))
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"variant": variant, "log_n": k, "ms": e0.elapsed_time(e1) / reps, "passes_ms": zk.capi.get_profile().get("passes")}), flush=True)


if __name__ == "__main__":
    main()
