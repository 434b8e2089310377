"""Point-range sharding of one large G1 MSM over the GPUs of a node.

MSM is linear in the point set, so rank r keeps the slice [r*n/G, (r+1)*n/G) of the SRS table
resident, receives the matching slice of the scalars, runs the whole Pippenger pipeline locally
down to one un-normalised point, and the G partial points (192 bytes each, extended Jacobian) are
stored into rank 0's HBM over NVLink from inside the MSM's last kernel and folded by whichever rank arrives last
(``b200zk_msm_g1_xchg_dev``).  One process per GPU; ``torch.distributed`` is the plumbing (process group, one 64-byte
broadcast of the IPC handle; NCCL on GPUs, gloo in the CPU tests of the host logic).  The same sharding driven by ONE
process over all GPUs is behind the plain C ABI (``b200zk_init_devices`` + ``b200zk_msm_g1``).

Batches of independent polynomials are split by :func:`split_batch` with no collective at all.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

from . import capi
from .capi import check, lib


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [start, end) of n points owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def split_batch(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent polynomials / proofs to ranks (no exchange needed)."""
    return list(range(rank, n_items, world))


class ShardedMSM:
    """One rank's share of a point-range sharded MSM (one process per GPU).

    ``bases_handle`` must hold exactly this rank's slice of the table (see :func:`shard_range`).

    ``exchange="peer"`` (default): the partial sums meet in rank 0's HBM through peer stores issued by the last
    kernel of every rank's MSM (``b200zk_msm_g1_xchg_dev``; the area is shared with CUDA IPC, its 64-byte handle is
    broadcast once over ``torch.distributed``); the rank that arrives last folds and normalises, no collective runs on
    the data path.  ``exchange="nccl"``: the plain all-gather of the 192-byte partials + a fold launch, kept as the
    comparison baseline."""

    def __init__(self, bases_handle: int, n_local: int, rank: int = 0, world: int = 1, group=None, exchange: str = "peer"):
        import torch

        self.torch = torch
        self.h = bases_handle
        self.n_local = n_local
        self.rank, self.world, self.group = rank, world, group
        self.exchange = exchange if world > 1 else "none"
        self.d_part = torch.zeros(192, dtype=torch.uint8, device="cuda")           # XYZZ partial of this rank
        self.d_all = torch.zeros(192 * world, dtype=torch.uint8, device="cuda")
        self.d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
        self.d_scalars = None
        self.copy_stream = None
        self.xchg = None
        if self.exchange == "peer":
            hbuf = torch.zeros(64, dtype=torch.uint8)
            xh = C.c_uint64(0)
            if rank == 0:
                raw = C.create_string_buffer(64)
                check(lib().b200zk_xchg_create(world, capi.addr(raw), C.byref(xh)))
                hbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
            dev = hbuf.cuda()
            torch.distributed.broadcast(dev, src=0, group=group)      # plumbing: 64 bytes, once
            if rank != 0:
                raw = bytes(dev.cpu().numpy())
                check(lib().b200zk_xchg_open(capi.addr(raw), world, rank, C.byref(xh)))
            self.xchg = xh.value

    def close(self) -> None:
        if self.xchg:
            self.torch.cuda.synchronize()
            if self.world > 1:
                self.torch.distributed.barrier(group=self.group)     # nobody unmaps while a peer may still store
            if self.rank != 0:
                check(lib().b200zk_xchg_close(self.xchg))
            if self.world > 1:
                self.torch.distributed.barrier(group=self.group)
            if self.rank == 0:
                check(lib().b200zk_xchg_close(self.xchg))
            self.xchg = None

    def timed_out(self) -> bool:
        if not self.xchg:
            return False
        v = C.c_uint32(0)
        check(lib().b200zk_xchg_status(self.xchg, C.byref(v)))
        return bool(v.value)

    def run_device(self, d_scalars, scalar_fmt: int = capi.FMT_CANONICAL, n: Optional[int] = None, offset: int = 0):
        """Scalars already in HBM -> result (canonical affine, 96 bytes) in ``self.d_out`` on every rank; asynchronous
        on torch's current stream."""
        torch = self.torch
        n = self.n_local if n is None else n
        st = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            check(lib().b200zk_msm_g1_dev(self.h, offset, d_scalars.data_ptr(), n, 1, scalar_fmt, 0, self.d_out.data_ptr(), st))
            return self.d_out
        if self.exchange == "peer":
            check(lib().b200zk_msm_g1_xchg_dev(self.h, offset, d_scalars.data_ptr(), n, scalar_fmt, self.xchg, 0,
                                               self.d_out.data_ptr(), st))
            return self.d_out
        check(lib().b200zk_msm_g1_partial_dev(self.h, offset, d_scalars.data_ptr(), n, scalar_fmt, self.d_part.data_ptr(), st))
        torch.distributed.all_gather_into_tensor(self.d_all, self.d_part, group=self.group)
        check(lib().b200zk_g1_sum_partials_dev(self.d_all.data_ptr(), self.world, 0, self.d_out.data_ptr(), st))
        return self.d_out

    def run_host(self, h_scalars_pinned, h_out_pinned, scalar_fmt: int = capi.FMT_CANONICAL) -> None:
        """End-to-end call with host buffers: H2D of this rank's scalars, MSM, exchange, D2H of the
        96-byte result, synchronised on return."""
        torch = self.torch
        if self.exchange == "peer":
            check(lib().b200zk_msm_g1_xchg(self.h, 0, h_scalars_pinned.data_ptr(), self.n_local, scalar_fmt, self.xchg,
                                           h_out_pinned.data_ptr()))
            return
        if self.d_scalars is None or self.d_scalars.numel() != h_scalars_pinned.numel():
            self.d_scalars = torch.empty(h_scalars_pinned.numel(), dtype=torch.uint8, device="cuda")
        self.d_scalars.copy_(h_scalars_pinned, non_blocking=True)
        out = self.run_device(self.d_scalars, scalar_fmt)
        h_out_pinned.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()


def combine_partials_host(partials: List[bytes]) -> bytes:
    """Sum of per-rank partial points given as canonical wire bytes (uploads, sums on the GPU)."""
    import torch

    n = len(partials)
    pts = b"".join(partials)
    h = C.c_uint64(0)
    check(lib().b200zk_bases_register(capi.addr(pts), n, capi.FMT_CANONICAL, 96, C.byref(h)))
    ones = (1).to_bytes(32, "little") * n
    out = C.create_string_buffer(96)
    try:
        check(lib().b200zk_msm_g1(h.value, 0, capi.addr(ones), n, capi.FMT_CANONICAL, capi.addr(out)))
    finally:
        lib().b200zk_bases_release(h.value)
    return out.raw


# ------------------------------------------------------------------------------------------------
# One large NTT over the GPUs of a node.  The transform itself runs inside the library from ONE process
# (b200zk_init_devices + b200zk_ntt_fr / b200zk_ntt_fr_sharded_dev; csrc/b200zk.cu, ntt_sharded): four-step decomposition
# n = R * C,
#   X[k1 + R k2] = sum_{i2} (w^R)^(i2 k2) * w^(i2 k1) * sum_{i1} (w^C)^(i1 k1) x[i1 C + i2],
# GPU g transforming the columns i2 in [g C/G, (g+1) C/G), storing row k1 of the twiddled intermediate matrix straight into
# the HBM of the GPU that owns it (the one exchange, peer stores from inside the kernel), then every GPU transforming its R/G
# rows.  What lives here is the plan's host-side mirror -- shapes, block layouts, who owns what -- for callers of the
# resident form and for the CPU tests of the decomposition (tests/test_dist_gloo.py).
# ------------------------------------------------------------------------------------------------
class FourStepPlan:
    """Shapes and index maps of the multi-GPU transform (mirror of ``ntt_shard_shape`` / ``ntt_sharded`` in csrc/b200zk.cu)."""

    MAX_LOG_R = 11          # the column transforms run in one pass of the NTT kernel

    def __init__(self, log_n: int, parts: int):
        if parts < 2 or parts & (parts - 1):
            raise ValueError("a sharded transform needs 2, 4, 8 or 16 GPUs")
        log_g = parts.bit_length() - 1
        self.log_n, self.parts = log_n, parts
        self.log_r = min(self.MAX_LOG_R, log_n - log_g)
        self.log_c = log_n - self.log_r
        self.R, self.C = 1 << self.log_r, 1 << self.log_c
        if self.R < parts or self.C < parts:
            raise ValueError("transform too small to shard over %d GPUs" % parts)
        self.Rl, self.Cl = self.R // parts, self.C // parts

    def block_in(self, x: Sequence, g: int) -> List[List]:
        """GPU g's input block [R][C/G] of a natural-order vector: x[i1 * C + g * Cl + u]."""
        return [[x[i1 * self.C + g * self.Cl + u] for u in range(self.Cl)] for i1 in range(self.R)]

    def twiddle_exponent(self, g: int, u: int, k1: int) -> int:
        """The column pass of GPU g multiplies output k1 of its column u by w^this."""
        return (g * self.Cl + u) * k1

    def row_owner(self, k1: int) -> Tuple[int, int]:
        """(GPU that runs the row transform of row k1, row index inside its block)."""
        return k1 // self.Rl, k1 % self.Rl

    def natural_from_blocks_out(self, blocks: Sequence[Sequence[Sequence]]) -> List:
        """Natural-order result from the per-GPU output blocks [C][R/G]: block h holds X[h * Rl + k1 + R * k2] at [k2][k1]."""
        out = [None] * (1 << self.log_n)
        for h, blk in enumerate(blocks):
            for k2 in range(self.C):
                for k1 in range(self.Rl):
                    out[h * self.Rl + k1 + self.R * k2] = blk[k2][k1]
        return out
