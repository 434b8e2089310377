"""b200zk -- B200 (sm_100a) backend for the BLS12-381 G1 MSM / Fr NTT hot path of the Halo2/KZG
prover and verifier used by input-output-hk/plutus-halo2-verifier-gen.

This package is plumbing over ``libb200zk.so`` (CUDA, C ABI declared in ``include/b200zk.h``):
 * :mod:`.capi`  -- ctypes binding of every exported symbol, error mapping;
 * :mod:`.host`  -- host-side mirror of the reference-facing interface for this path
   (``ParamsKZG`` residency, ``KZGCommitmentScheme.commit/commit_lagrange``,
   ``EvaluationDomain`` transforms, ``DualMSM.eval``), with the reference's argument meaning;
 * :mod:`.dist`  -- point-range sharding of one large MSM over the GPUs of a node
   (one process per GPU, ``torch.distributed``).

There is no CPU fallback: importing works anywhere, but every compute entry point raises
:class:`B200zkError` unless the CUDA library loads and a B200 is present.
"""
from .capi import (  # noqa: F401
    B200zkError,
    FMT_CANONICAL,
    FMT_MONT,
    NTT_COSET_IN,
    NTT_COSET_OUT,
    NTT_INVERSE_SCALE,
    NTT_MONT,
    lib,
    lib_path,
    init,
    shutdown,
    device_info,
    launch_count,
)
from . import host  # noqa: F401

__all__ = [
    "B200zkError", "FMT_CANONICAL", "FMT_MONT", "NTT_COSET_IN", "NTT_COSET_OUT", "NTT_INVERSE_SCALE",
    "NTT_MONT", "lib", "lib_path", "init", "shutdown", "device_info", "launch_count", "host",
]
