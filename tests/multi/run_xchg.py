"""One process per GPU: point-range sharded MSM with the fused peer-store exchange (dist.ShardedMSM, exchange="peer")
against the CPU oracle, plus the NCCL all-gather baseline.  Launched by tests/test_gpu_multi.py under torch.distributed.run."""
import ctypes as C
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import Oracle, _build_oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    zk.init(local)
    L = zk.lib()
    orc = Oracle(_build_oracle())
    n = (1 << 15) + 11
    pts = orc.synth_bases(0xB200, 0, n)
    sc = orc.synth_scalars(3, 0, n)
    want = orc.msm(pts, sc, n)
    s, e = zdist.shard_range(n, rank, world)
    h = C.c_uint64(0)
    zk.capi.check(L.b200zk_bases_register(zk.capi.addr(pts[96 * s:96 * e]), e - s, 0, 96, C.byref(h)))
    d_sc = torch.frombuffer(bytearray(sc[32 * s:32 * e]), dtype=torch.uint8).cuda()
    checks = {}
    for mode in ("peer", "nccl"):
        m = zdist.ShardedMSM(h.value, e - s, rank, world, exchange=mode)
        for rep in range(3):                       # sequence numbers advance, counters are never reset
            out = m.run_device(d_sc)
            torch.cuda.synchronize()
            checks["%s device rep %d" % (mode, rep)] = bytes(out.cpu().numpy()) == want
        h_sc = torch.frombuffer(bytearray(sc[32 * s:32 * e]), dtype=torch.uint8).pin_memory()
        h_out = torch.zeros(96, dtype=torch.uint8).pin_memory()
        m.run_host(h_sc, h_out)
        checks["%s host" % mode] = bytes(h_out.numpy()) == want
        # a shorter slice: ranks with nothing to add still arrive at the exchange (with the identity)
        k0 = min(100, n // world)
        out = m.run_device(d_sc, n=k0 if rank == 0 else 0)
        torch.cuda.synchronize()
        checks["%s short" % mode] = bytes(out.cpu().numpy()) == orc.msm(pts[:96 * k0], sc[:32 * k0], k0)
        checks["%s no timeout" % mode] = not m.timed_out()
        m.close()
    dist.barrier()
    res = json.dumps({"rank": rank, "ok": all(checks.values()), "checks": checks})
    outdir = os.environ.get("XCHG_OUT")
    if outdir:                                     # one file per rank: the ranks' stdout lines may interleave
        with open(os.path.join(outdir, "rank%d.json" % rank), "w") as f:
            f.write(res)
    print(res, flush=True)
    dist.destroy_process_group()
    return 0 if all(checks.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
