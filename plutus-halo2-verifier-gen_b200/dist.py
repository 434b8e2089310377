"""Point-range sharding of one large G1 MSM over the GPUs of a node.

MSM is linear in the point set, so rank r keeps the slice [r*n/G, (r+1)*n/G) of the SRS table
resident, receives the matching slice of the scalars, runs the whole Pippenger pipeline locally
down to one un-normalised point, and the G partial points (192 bytes each, extended Jacobian) are
exchanged with one all-gather over NVLink and folded on the device (``b200zk_g1_sum_partials_dev``).  One process per GPU;
``torch.distributed`` is the plumbing (NCCL on GPUs, gloo in the CPU tests of the host logic).

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
    """One rank's share of a point-range sharded MSM.

    ``bases_handle`` must hold exactly this rank's slice of the table (see :func:`shard_range`)."""

    def __init__(self, bases_handle: int, n_local: int, rank: int = 0, world: int = 1, group=None):
        import torch

        self.torch = torch
        self.h = bases_handle
        self.n_local = n_local
        self.rank, self.world, self.group = rank, world, group
        self.d_part = torch.zeros(192, dtype=torch.uint8, device="cuda")           # XYZZ partial of this rank
        self.d_all = torch.zeros(192 * world, dtype=torch.uint8, device="cuda")
        self.d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
        self.d_scalars = None

    def run_device(self, d_scalars, scalar_fmt: int = capi.FMT_CANONICAL):
        """Scalars already in HBM -> result (canonical affine, 96 bytes) in ``self.d_out``; asynchronous
        on torch's current stream."""
        torch = self.torch
        st = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            check(lib().b200zk_msm_g1_dev(self.h, 0, d_scalars.data_ptr(), self.n_local, 1, scalar_fmt, 0,
                                          self.d_out.data_ptr(), st))
            return self.d_out
        check(lib().b200zk_msm_g1_partial_dev(self.h, 0, d_scalars.data_ptr(), self.n_local, scalar_fmt,
                                              self.d_part.data_ptr(), st))
        torch.distributed.all_gather_into_tensor(self.d_all, self.d_part, group=self.group)
        check(lib().b200zk_g1_sum_partials_dev(self.d_all.data_ptr(), self.world, 0, self.d_out.data_ptr(), st))
        return self.d_out

    def run_host(self, h_scalars_pinned, h_out_pinned, scalar_fmt: int = capi.FMT_CANONICAL) -> None:
        """End-to-end call with host buffers: H2D of this rank's scalars, MSM, exchange, D2H of the
        96-byte result, synchronised on return."""
        torch = self.torch
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
