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


class _CpuNttOps:
    """Stands in for the per-GPU kernels in the gloo test of the sharded NTT: the checker's transforms on CPU tensors."""

    def __init__(self, pyref):
        self.P = pyref

    @staticmethod
    def _ints(t):
        b = bytes(t.contiguous().view(-1).numpy())
        return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(len(b) // 32)]

    @staticmethod
    def _store(t, vals):
        import torch
        raw = b"".join(v.to_bytes(32, "little") for v in vals)
        t.view(-1).copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))

    def ntt_batch(self, t, batch, log_len, omega, inverse):
        vals, n = self._ints(t), 1 << log_len
        out = []
        for b in range(batch):
            r = self.P.ntt(vals[b * n:(b + 1) * n], omega)
            if inverse:
                ninv = self.P.fr_inv(n)
                r = [v * ninv % self.P.R_MOD for v in r]
            out += r
        self._store(t, out)

    def power_table(self, base, row0, rows, cols, device):
        import torch
        t = torch.empty(rows * cols * 32, dtype=torch.uint8)
        self._store(t, [pow(base, (row0 + r) * c, self.P.R_MOD) for r in range(rows) for c in range(cols)])
        return t

    def mul_table(self, t, table):
        self._store(t, [a * b % self.P.R_MOD for a, b in zip(self._ints(t), self._ints(table))])


def _ntt_worker(rank, world, port, log_n, inverse, q):
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
    full = [P.splitmix64(1000 + i) * P.splitmix64(7 + i) % P.R_MOD for i in range(n)]
    s, e = zdist.shard_range(n, rank, world)
    mine = torch.frombuffer(bytearray(b"".join(v.to_bytes(32, "little") for v in full[s:e])), dtype=torch.uint8)
    w = P.omega(log_n)
    if inverse:
        w = P.fr_inv(w)
    plan = zdist.ShardedNTT(log_n, w, rank, world, inverse=inverse, ops=_CpuNttOps(P))
    out = plan.run(mine)
    q.put((rank, bytes(out.numpy())))
    dist.destroy_process_group()


@pytest.mark.parametrize("log_n,inverse,world", [(6, False, 2), (7, False, 2), (7, True, 2), (8, False, 4)])
def test_sharded_ntt_world2(pyref, log_n, inverse, world):
    """The exchange / transpose / twiddle logic of dist.ShardedNTT on two (and four) gloo ranks, with the checker's
    transforms standing in for the kernels: the concatenated result equals the checker's transform of the whole vector."""
    import torch.multiprocessing as mp
    port = 29617 + (os.getpid() + log_n + 7 * inverse + 13 * world) % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ntt_worker, args=(r, world, port, log_n, inverse, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P = pyref
    n = 1 << log_n
    full = [P.splitmix64(1000 + i) * P.splitmix64(7 + i) % P.R_MOD for i in range(n)]
    w = P.omega(log_n)
    want = P.intt(full, w) if inverse else P.ntt(full, w)
    raw = b"".join(got[r] for r in range(world))
    assert [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(n)] == want


def test_sharded_ntt_single_rank_matches(pyref):
    """world = 1 runs the same four-step path without a process group."""
    import torch
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
    P = pyref
    log_n = 6
    n = 1 << log_n
    full = [P.splitmix64(5 + i) % P.R_MOD for i in range(n)]
    t = torch.frombuffer(bytearray(b"".join(v.to_bytes(32, "little") for v in full)), dtype=torch.uint8)
    out = zdist.ShardedNTT(log_n, P.omega(log_n), 0, 1, ops=_CpuNttOps(P)).run(t)
    raw = bytes(out.numpy())
    assert [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(n)] == P.ntt(full, P.omega(log_n))
