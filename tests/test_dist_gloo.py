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


def _four_step_worker(rank, world, port, log_n, q):
    """One rank of the multi-GPU transform with the checker's transforms standing in for the kernels and gloo's all_to_all for the
    peer stores: column block in, column transforms with root w^C, twiddles w^((g Cl + u) k1), every row k1 sent to its owner,
    row transforms with root w^R, transposed output block."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyref as P
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1 << log_n
    x = [P.splitmix64(17 * i + 3) % P.R_MOD for i in range(n)]
    w = P.omega(log_n)
    plan = zdist.FourStepPlan(log_n, world)
    blk = plan.block_in(x, rank)                                   # [R][Cl]
    wR, wC = pow(w, plan.R, P.R_MOD), pow(w, plan.C, P.R_MOD)
    cols = []
    for u in range(plan.Cl):                                       # column transforms + twiddles
        y = P.ntt([blk[i1][u] for i1 in range(plan.R)], wC)
        cols.append([y[k1] * pow(w, plan.twiddle_exponent(rank, u, k1), P.R_MOD) % P.R_MOD for k1 in range(plan.R)])
    # the exchange: rows k1 of block h go to rank h, as [Rl][Cl] (what the column pass stores into its peer's HBM)
    def pack(vals):
        return torch.frombuffer(bytearray(b"".join(v.to_bytes(32, "little") for v in vals)), dtype=torch.uint8)
    send = [pack([cols[u][h * plan.Rl + r] for r in range(plan.Rl) for u in range(plan.Cl)]) for h in range(world)]
    # (gloo has no all_to_all: everybody gathers everybody's send list and keeps the piece addressed to it)
    mine = torch.cat(send)
    everyone = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)
    seg = send[0].numel()
    recv = [everyone[g][rank * seg:(rank + 1) * seg] for g in range(world)]
    rows = [[0] * plan.C for _ in range(plan.Rl)]
    for g in range(world):
        b = bytes(recv[g].numpy())
        for r in range(plan.Rl):
            for u in range(plan.Cl):
                o = 32 * (r * plan.Cl + u)
                rows[r][g * plan.Cl + u] = int.from_bytes(b[o:o + 32], "little")
        assert all(plan.row_owner(rank * plan.Rl + r) == (rank, r) for r in range(plan.Rl))
    z = [P.ntt(rows[r], wR) for r in range(plan.Rl)]               # row transforms
    out_blk = [[z[r][k2] for r in range(plan.Rl)] for k2 in range(plan.C)]   # transposed store: [C][Rl]
    q.put((rank, out_blk))
    dist.destroy_process_group()


@pytest.mark.parametrize("log_n,world", [(6, 2), (9, 2), (8, 4)])
def test_four_step_plan_over_gloo_ranks(pyref, zk, log_n, world):
    """The decomposition the multi-GPU transform implements (csrc/b200zk.cu, ntt_sharded; mirrored by dist.FourStepPlan): block
    layouts, twiddle exponents, row ownership and the transposed output reproduce the plain transform."""
    import torch.multiprocessing as mp
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    port = 29700 + os.getpid() % 200 + log_n
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_four_step_worker, args=(r, world, port, log_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plan = zdist.FourStepPlan(log_n, world)
    n = 1 << log_n
    x = [pyref.splitmix64(17 * i + 3) % pyref.R_MOD for i in range(n)]
    assert plan.natural_from_blocks_out([got[r] for r in range(world)]) == pyref.ntt(x, pyref.omega(log_n))


def test_four_step_plan_shapes(zk):
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    p = zdist.FourStepPlan(24, 8)
    assert (p.log_r, p.log_c, p.Cl, p.Rl) == (11, 13, 1024, 256)
    p = zdist.FourStepPlan(13, 8)
    assert (p.log_r, p.log_c, p.Cl, p.Rl) == (10, 3, 1, 128)
    with pytest.raises(ValueError):
        zdist.FourStepPlan(3, 8)
    with pytest.raises(ValueError):
        zdist.FourStepPlan(20, 3)
