"""ctypes binding of libb200zk.so (every symbol of include/b200zk.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make``; if it is missing or
fails to load, :func:`lib` raises -- there is no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

FMT_CANONICAL = 0
FMT_MONT = 1
BASES_NO_WINDOW_TABLES = 0x100
BASES_SHARD = 0x200
BASES_REPLICATE = 0x400
NTT_INVERSE_SCALE = 1
NTT_COSET_IN = 2
NTT_COSET_OUT = 4
NTT_MONT = 8

ERR_NAMES = {
    -1: "INVALID_ARG", -2: "CUDA", -3: "NO_DEVICE", -4: "BAD_HANDLE", -5: "OOM", -6: "NOT_INIT", -7: "BAD_POINT",
}

_HERE = os.path.dirname(os.path.abspath(__file__))
_lock = threading.Lock()
_lib = None


class B200zkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("b200zk error %d (%s): %s" % (code, ERR_NAMES.get(code, "?"), msg))
        self.code = code


def lib_path() -> str:
    return os.environ.get("B200ZK_LIB", os.path.join(_HERE, "libb200zk.so"))


# name -> (restype, argtypes); kept in one table so the "exports every declared symbol" test
# can walk it against include/b200zk.h
_u8p = C.c_void_p  # byte buffers are passed as raw addresses (bytes, bytearray, numpy, torch all work)
SIGNATURES = {
    "b200zk_init_devices": (C.c_int32, [C.c_void_p, C.c_int32]),
    "b200zk_init": (C.c_int32, [C.c_int32]),
    "b200zk_device_count": (C.c_int32, []),
    "b200zk_set_device": (C.c_int32, [C.c_int32]),
    "b200zk_shutdown": (C.c_int32, []),
    "b200zk_last_error": (C.c_int32, [C.c_char_p, C.c_size_t]),
    "b200zk_device_info": (C.c_int32, [C.c_char_p, C.c_size_t]),
    "b200zk_host_alloc": (C.c_int32, [C.POINTER(C.c_void_p), C.c_size_t]),
    "b200zk_host_free": (C.c_int32, [C.c_void_p]),
    "b200zk_bases_register": (C.c_int32, [_u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "b200zk_bases_register_dev": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "b200zk_bases_release": (C.c_int32, [C.c_uint64]),
    "b200zk_bases_read": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_uint64, _u8p]),
    "b200zk_msm_g1": (C.c_int32, [C.c_uint64, C.c_uint64, _u8p, C.c_uint64, C.c_uint32, _u8p]),
    "b200zk_msm_g1_batch": (C.c_int32, [C.c_uint64, C.c_uint64, _u8p, C.c_uint64, C.c_uint32, C.c_uint32, _u8p]),
    "b200zk_msm_g1_batch_ptrs": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, _u8p]),
    "b200zk_bases_layout": (C.c_int32, [C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_uint32)]),
    "b200zk_msm_g1_sharded_dev": (C.c_int32, [C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, _u8p]),
    "b200zk_msm_g1_adhoc": (C.c_int32, [_u8p, C.c_uint32, _u8p, C.c_uint32, C.c_uint64, _u8p]),
    "b200zk_msm_g1_dev": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200zk_msm_g1_partial_dev": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "b200zk_g1_sum_partials_dev": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200zk_xchg_create": (C.c_int32, [C.c_uint32, _u8p, C.POINTER(C.c_uint64)]),
    "b200zk_xchg_open": (C.c_int32, [_u8p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "b200zk_xchg_close": (C.c_int32, [C.c_uint64]),
    "b200zk_xchg_status": (C.c_int32, [C.c_uint64, C.POINTER(C.c_uint32)]),
    "b200zk_msm_g1_xchg_dev": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "b200zk_msm_g1_xchg": (C.c_int32, [C.c_uint64, C.c_uint64, _u8p, C.c_uint64, C.c_uint32, C.c_uint64, _u8p]),
    "b200zk_g1_sum_dev": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200zk_ntt_fr": (C.c_int32, [_u8p, C.c_uint32, _u8p, C.c_uint32, _u8p]),
    "b200zk_ntt_fr_batch": (C.c_int32, [_u8p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, _u8p]),
    "b200zk_ntt_fr_batch_ptrs": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, _u8p]),
    "b200zk_ntt_fr_dev": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, _u8p, C.c_void_p]),
    "b200zk_ntt_sharded_layout": (C.c_int32, [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b200zk_ntt_fr_sharded_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, _u8p]),
    "b200zk_g1_compress": (C.c_int32, [_u8p, _u8p]),
    "b200zk_dev_alloc": (C.c_int32, [C.POINTER(C.c_void_p), C.c_size_t]),
    "b200zk_dev_free": (C.c_int32, [C.c_void_p]),
    "b200zk_dev_upload": (C.c_int32, [C.c_void_p, _u8p, C.c_size_t]),
    "b200zk_dev_download": (C.c_int32, [_u8p, C.c_void_p, C.c_size_t]),
    "b200zk_g1_decompress_batch": (C.c_int32, [_u8p, C.c_uint64, _u8p, C.c_void_p]),
    "b200zk_g1_decompress_dev": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200zk_srs_generate_dev": (C.c_int32, [_u8p, C.c_uint32, _u8p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200zk_g1_fixed_mul_dev": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "b200zk_g1_export_dev": (C.c_int32, [C.c_void_p, C.c_uint64, _u8p]),
    "b200zk_fr_convert_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    "b200zk_fr_power_table_dev": (C.c_int32, [_u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "b200zk_fr_extend_dev": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    "b200zk_fr_pointwise_dev": (C.c_int32, [C.c_uint32, C.c_void_p, C.c_void_p, _u8p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "b200zk_fr_lincomb_dev": (C.c_int32, [C.c_void_p, _u8p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]),
    "b200zk_fr_batch_invert_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "b200zk_fr_running_product_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, _u8p, C.c_uint32, C.c_void_p]),
    "b200zk_fr_kate_div_dev": (C.c_int32, [C.c_void_p, C.c_uint64, _u8p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200zk_gate_program_create": (C.c_int32, [C.c_void_p, C.c_uint32, _u8p, C.c_uint32, C.c_void_p, C.c_uint32, _u8p, C.c_uint32,
                                               C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "b200zk_gate_program_set_const": (C.c_int32, [C.c_uint64, C.c_uint32, _u8p]),
    "b200zk_gate_program_run_dev": (C.c_int32, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "b200zk_gate_program_release": (C.c_int32, [C.c_uint64]),
    "b200zk_transcript_new": (C.c_int32, [C.POINTER(C.c_uint64)]),
    "b200zk_transcript_free": (C.c_int32, [C.c_uint64]),
    "b200zk_transcript_common_scalar": (C.c_int32, [C.c_uint64, _u8p]),
    "b200zk_transcript_common_point": (C.c_int32, [C.c_uint64, _u8p]),
    "b200zk_transcript_squeeze": (C.c_int32, [C.c_uint64, _u8p]),
    "b200zk_h2mo_open_dev": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, _u8p, C.c_uint32, _u8p,
                                         C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200zk_h2mo_prepare": (C.c_int32, [C.c_uint64, _u8p, C.c_uint32, C.c_void_p, _u8p, _u8p, C.c_uint32, _u8p, C.c_size_t,
                                        C.POINTER(C.c_uint64), _u8p]),
    "b200zk_h2mo_scalars": (C.c_int32, [C.c_uint32, C.c_void_p, _u8p, _u8p, C.c_uint32, _u8p, _u8p, C.c_uint32, _u8p, C.c_size_t, _u8p, _u8p]),
    "b200zk_guard_eval": (C.c_int32, [C.c_void_p, C.c_uint32, _u8p, _u8p, _u8p]),
    "b200zk_guard_free": (C.c_int32, [C.c_uint64]),
    "b200zk_g1_synth_bases_dev": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "b200zk_selftest_field": (C.c_int32, [C.c_uint32, C.c_uint32, _u8p, _u8p, _u8p, C.c_uint64]),
    "b200zk_microbench": (C.c_int32, [C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200zk_launch_count": (C.c_uint64, []),
    "b200zk_set_msm_tuning": (C.c_int32, [C.c_uint32, C.c_uint32]),
    "b200zk_set_profiling": (C.c_int32, [C.c_uint32]),
    "b200zk_get_profile": (C.c_int32, [C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.c_uint32, C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}


def lib() -> C.CDLL:
    """Loads libb200zk.so once.  Raises B200zkError if the CUDA extension is missing."""
    global _lib
    with _lock:
        if _lib is None:
            path = lib_path()
            if not os.path.exists(path):
                raise B200zkError(-3, "CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; "
                                      "g.build()'` (or `make`); there is no CPU fallback" % path)
            try:
                handle = C.CDLL(path)
            except OSError as e:  # pragma: no cover - depends on the box
                raise B200zkError(-3, "cannot load %s: %s" % (path, e))
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().b200zk_last_error(buf, len(buf))
    return buf.value.decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise B200zkError(rc, last_error())


def addr(buf) -> int:
    """Raw address of a bytes / bytearray / ctypes / numpy / torch buffer (None -> 0)."""
    if buf is None:
        return 0
    if isinstance(buf, int):
        return buf
    if isinstance(buf, (bytes, bytearray)):
        return C.addressof((C.c_char * len(buf)).from_buffer(buf)) if isinstance(buf, bytearray) else \
            C.cast(C.c_char_p(buf), C.c_void_p).value
    if hasattr(buf, "data_ptr"):       # torch tensor
        return buf.data_ptr()
    if hasattr(buf, "ctypes"):         # numpy array
        return buf.ctypes.data
    return C.addressof(buf)            # ctypes array


def init(device: int = -1) -> None:
    check(lib().b200zk_init(device))


def init_devices(device_ids=None, n: int = 0) -> None:
    """Binds several GPUs to this process (b200zk_init_devices): ``device_ids`` CUDA ordinals, or the first ``n``."""
    if device_ids is None:
        check(lib().b200zk_init_devices(None, n))
    else:
        arr = (C.c_int32 * len(device_ids))(*device_ids)
        check(lib().b200zk_init_devices(C.addressof(arr), len(device_ids)))


def bases_layout(handle: int):
    """[(device index, start, n), ...] and whether the table is replicated."""
    cnt, rep = C.c_uint32(0), C.c_uint32(0)
    dev = (C.c_int32 * 16)()
    start = (C.c_uint64 * 16)()
    num = (C.c_uint64 * 16)()
    check(lib().b200zk_bases_layout(handle, 16, C.byref(cnt), C.addressof(dev), C.addressof(start), C.addressof(num), C.byref(rep)))
    return [(dev[i], start[i], num[i]) for i in range(cnt.value)], bool(rep.value)


def device_count() -> int:
    return int(lib().b200zk_device_count())


def set_device(index: int) -> None:
    check(lib().b200zk_set_device(index))


def shutdown() -> None:
    check(lib().b200zk_shutdown())


def device_info() -> str:
    buf = C.create_string_buffer(256)
    check(lib().b200zk_device_info(buf, len(buf)))
    return buf.value.decode()


def launch_count() -> int:
    return int(lib().b200zk_launch_count())


def set_profiling(enable: bool) -> None:
    check(lib().b200zk_set_profiling(1 if enable else 0))


def get_profile() -> dict:
    """Phase times (ms) of the last MSM / NTT call, measured with CUDA events on its stream."""
    kind, n, c, w = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    ms = (C.c_double * 8)()
    check(lib().b200zk_get_profile(C.byref(kind), ms, 8, C.byref(n), C.byref(c), C.byref(w)))
    phases = [ms[i] for i in range(n.value)]
    if kind.value == 1:
        names = ["sort", "accumulate", "tail"]
        return {"kind": "msm", "window_bits": c.value, "windows": w.value, **dict(zip(names, phases))}
    return {"kind": "ntt", "passes": phases}
