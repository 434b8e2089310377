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
from typing import List, Optional, Tuple

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
# One large NTT over the GPUs of a node (worth it at 2^23..2^24 because the transform is
# bound by the integer pipe, not by the exchange): four-step decomposition n = R * C,
#   X[k1 + R k2] = sum_{i2} (w^R)^(i2 k2) * w^(i2 k1) * sum_{i1} (w^C)^(i1 k1) x[i1 C + i2],
# natural order in and out, block-distributed (rank g holds elements [g n/G, (g+1) n/G)).
#   a2a #1 : row blocks -> column blocks          (each rank gets all i1 for its i2 block)
#   local  : C/G transforms of length R, times the twiddle block w^(i2 k1)   (b200zk_ntt_fr_dev, _fr_pointwise_dev)
#   a2a #2 : -> all i2 for the rank's k1 block
#   local  : R/G transforms of length C
#   a2a #3 : back to natural order (k = k2 R + k1: the rank's k2 block, all k1)
# Each exchange moves n/G elements per rank over NVLink (64 MiB at 2^24 on 8 GPUs); packing is a strided copy.
# ------------------------------------------------------------------------------------------------
class GpuNttOps:
    """The per-rank primitives of :class:`ShardedNTT` on libb200zk.so (device tensors, torch's current stream)."""

    def __init__(self):
        import torch

        self.torch = torch

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def ntt_batch(self, t, batch: int, log_len: int, omega: int, inverse: bool) -> None:
        from .host import fr_bytes

        om = fr_bytes(omega)
        flags = capi.NTT_INVERSE_SCALE if inverse else 0
        check(lib().b200zk_ntt_fr_dev(t.data_ptr(), batch, log_len, capi.addr(om), flags, None, self._stream()))

    def power_table(self, base: int, row0: int, rows: int, cols: int, device):
        from .host import fr_bytes

        t = self.torch.empty(rows * cols * 32, dtype=self.torch.uint8, device=device)
        bb = fr_bytes(base)
        check(lib().b200zk_fr_power_table_dev(capi.addr(bb), row0, rows, cols, t.data_ptr(), self._stream()))
        return t

    def mul_table(self, t, table) -> None:
        """t[i] *= table[i]; t canonical, table in Montgomery form (so the product is canonical again)."""
        check(lib().b200zk_fr_pointwise_dev(0, t.data_ptr(), table.data_ptr(), None, t.data_ptr(), t.numel() // 32, self._stream()))


class ShardedNTT:
    """Rank-local state of one forward or inverse NTT of 2^log_n elements over `world` ranks."""

    def __init__(self, log_n: int, omega: int, rank: int, world: int, group=None, inverse: bool = False, ops=None, modulus: Optional[int] = None):
        import torch

        from .host import R_MOD

        self.torch = torch
        self.p = modulus or R_MOD
        if world & (world - 1):
            raise ValueError("world size must be a power of two")
        log_g = world.bit_length() - 1
        self.log_r = max(log_n // 2, log_g)
        self.log_c = log_n - self.log_r
        if self.log_c < log_g:
            raise ValueError("transform too small for this many ranks")
        self.log_n, self.omega, self.rank, self.world, self.group, self.inverse = log_n, omega % self.p, rank, world, group, inverse
        self.ops = ops or GpuNttOps()
        self.R, self.Cc = 1 << self.log_r, 1 << self.log_c
        self.Rl, self.Cl = self.R // world, self.Cc // world
        self.table = None

    def _a2a(self, send):
        recv = self.torch.empty_like(send)
        if self.world == 1:
            recv.copy_(send)
        else:
            self.torch.distributed.all_to_all_single(recv, send, group=self.group)
        return recv

    def run(self, x_local):
        """x_local: uint8 tensor of (n / world) * 32 bytes, this rank's contiguous slice in natural order (canonical
        or Montgomery Fr; the transform is linear).  Returns the rank's slice of the result, same layout."""
        G, R, Cc, Rl, Cl = self.world, self.R, self.Cc, self.Rl, self.Cl
        w = self.omega
        if self.table is None:   # w^((rank*Cl + u) * k1), laid out like the data it multiplies: [u][k1]
            self.table = self.ops.power_table(w, self.rank * Cl, Cl, R, x_local.device)
        # 32-byte elements are moved as pairs of 16-byte words (complex128 is only a container type here)
        W16 = self.torch.complex128
        x = x_local.view(W16).view(Rl, G, Cl, 2)
        # a2a #1: [Rl][G][Cl] -> send [G][Rl][Cl]; receive rows of every source rank: [R][Cl]
        a = self._a2a(x.permute(1, 0, 2, 3).contiguous())
        # local: [R][Cl] -> [Cl][R], transforms over i1 with root w^C, then the twiddle block
        y = a.view(R, Cl, 2).permute(1, 0, 2).contiguous()
        self.ops.ntt_batch(y.view(-1).view(self.torch.uint8), Cl, self.log_r, pow(w, Cc, self.p), self.inverse)
        self.ops.mul_table(y.view(-1).view(self.torch.uint8), self.table)
        # a2a #2: [Cl][G][Rl] -> send [G][Cl][Rl]; receive [C][Rl]
        b = self._a2a(y.view(Cl, G, Rl, 2).permute(1, 0, 2, 3).contiguous())
        # local: [C][Rl] -> [Rl][C], transforms over i2 with root w^R
        z = b.view(Cc, Rl, 2).permute(1, 0, 2).contiguous()
        self.ops.ntt_batch(z.view(-1).view(self.torch.uint8), Rl, self.log_c, pow(w, R, self.p), self.inverse)
        # a2a #3: z[k1_local][k2] -> natural order k = k2 R + k1: send [G][Rl][Cl], receive [R][Cl] -> [Cl][R]
        c = self._a2a(z.view(Rl, G, Cl, 2).permute(1, 0, 2, 3).contiguous())
        return c.view(R, Cl, 2).permute(1, 0, 2).contiguous().view(-1).view(self.torch.uint8)
