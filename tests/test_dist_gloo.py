"""Host-side logic of the N > 1 path on CPU: world-size-2 `gloo` processes shard a point range,
compute their partial MSM with the checker (standing in for the per-GPU kernel), all-gather the
96-byte partial points and combine -- the same partition / exchange / combine steps that
dist.ShardedMSM runs over NCCL."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions(zk):
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    for n in (0, 1, 7, 8, 1000, (1 << 24), (1 << 24) + 5):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                s, e = zdist.shard_range(n, r, world)
                assert 0 <= s <= e <= n and (e - s) in (n // world, n // world + 1)
                cover.append((s, e))
            assert cover[0][0] == 0 and cover[-1][1] == n
            assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    assert zdist.split_batch(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((zdist.split_batch(23, r, 8) for r in range(8)), [])) == list(range(23))


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import Oracle, _build_oracle
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle(_build_oracle())
    s, e = zdist.shard_range(n, rank, world)
    bases = orc.synth_bases(0xB200, s, e - s)          # this rank's slice of the table
    scal = orc.synth_scalars(1, s, e - s)              # ... and of the scalars
    part = orc.msm(bases, scal, e - s)                 # stands in for b200zk_msm_g1_dev on this rank's GPU
    mine = torch.frombuffer(bytearray(part), dtype=torch.uint8)
    gathered = [torch.zeros(96, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, mine)
    total = orc.g1_sum(b"".join(bytes(t.numpy()) for t in gathered), world)
    q.put((rank, total))
    dist.destroy_process_group()


def test_point_range_sharding_world2(oracle):
    import torch.multiprocessing as mp
    n, world, port = 3001, 2, 29517 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oracle.msm(oracle.synth_bases(0xB200, 0, n), oracle.synth_scalars(1, 0, n), n)
    assert got[0] == got[1] == full
