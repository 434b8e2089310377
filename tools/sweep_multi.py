"""Multi-GPU part of the reporting grid: point-range sharded G1 MSM at several sizes and
scalar distributions on N GPUs of one node.  Launch with torchrun, one process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/sweep_multi.py --sizes 20,22,24 --dists U,S
Rank 0 prints one JSON line per cell (device-resident and end-to-end, max over ranks)."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
from circuit_bench import prover_like  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="20,22,24")
    ap.add_argument("--dists", default="U,S")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ntt-sizes", default="22,24")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    zk.init(local_rank)
    lib = zk.lib()
    st = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def mx(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for log_n in [int(x) for x in args.sizes.split(",")]:
        n = 1 << log_n
        s0, s1 = zdist.shard_range(n, rank, world)
        nl = s1 - s0
        d_b = torch.empty(96 * nl, dtype=torch.uint8, device="cuda")
        zk.capi.check(lib.b200zk_g1_synth_bases_dev(bench.BASE_SEED, s0, nl, d_b.data_ptr(), st))
        torch.cuda.synchronize()
        h = C.c_uint64(0)
        zk.capi.check(lib.b200zk_bases_register_dev(d_b.data_ptr(), nl, zk.FMT_MONT, 96, C.byref(h)))
        del d_b
        torch.cuda.empty_cache()
        msm = zdist.ShardedMSM(h.value, nl, rank, world)
        for name in args.dists.split(","):
            if name == "U":
                k = bench.synth_scalars_np(1, s0, nl)
            else:
                k = prover_like(np, 7 + rank, nl)
            h_sc = torch.from_numpy(np.ascontiguousarray(k).view(np.uint8).reshape(-1)).pin_memory()
            d_sc = h_sc.cuda()
            h_out = torch.zeros(96, dtype=torch.uint8).pin_memory()
            for _ in range(3):
                msm.run_device(d_sc)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                msm.run_device(d_sc)
            e1.record()
            barrier()
            ms = mx(e0.elapsed_time(e1)) / args.reps
            msm.run_host(h_sc, h_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                msm.run_host(h_sc, h_out)
            barrier()
            e2e = mx(time.perf_counter() - t0) / args.reps * 1e3
            if rank == 0:
                print(json.dumps({"op": "msm_g1_sharded", "n_gpus": world, "log_n": log_n, "dist": name, "ms": ms,
                                  "points_per_s": n / ms * 1e3, "e2e_ms": e2e, "e2e_points_per_s": n / e2e * 1e3,
                                  "imad_frac_48k_per_gpu": (n * bench.LMAC_PER_POINT / (ms * 1e-3)) / world / 9.22e12}), flush=True)
        zk.capi.check(lib.b200zk_bases_release(h.value))
    # ---- a batch of independent NTTs split by rank (no collective).  ONE transform over several GPUs runs from a single
    #      process inside the library now (b200zk_ntt_fr / b200zk_ntt_fr_sharded_dev; measured by bench.py's single_process leg)
    for log_n in [int(x) for x in args.ntt_sizes.split(",") if x]:
        n = 1 << log_n
        w = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - log_n), bench.R_MOD)
        wb = w.to_bytes(32, "little")
        items = zdist.split_batch(4 * world, rank, world)
        batch = torch.from_numpy(bench.synth_scalars_np(5 + rank, 0, n * len(items)).view(np.uint8).reshape(-1)).cuda()
        zk.capi.check(lib.b200zk_ntt_fr_dev(batch.data_ptr(), len(items), log_n, zk.capi.addr(wb), 0, 0, st))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            zk.capi.check(lib.b200zk_ntt_fr_dev(batch.data_ptr(), len(items), log_n, zk.capi.addr(wb), 0, 0, st))
        e1.record()
        barrier()
        msb = mx(e0.elapsed_time(e1)) / args.reps
        if rank == 0:
            print(json.dumps({"op": "ntt_fr_batch_split", "n_gpus": world, "log_n": log_n, "transforms": 4 * world, "ms": msb,
                              "elements_per_s": 4 * world * n / msb * 1e3}), flush=True)
        del batch
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
