#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 MSM / NTT backend.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the CPU path, timed on the host cores

Metric (BASELINE.json): BLS12-381 G1 MSM points/s on 2^24 synthetic points with uniform scalars
(distribution "U"); a step is one full MSM.  With N > 1 GPUs the same 2^24
points are sharded by point range (strong scaling) and the N partial points are all-gathered over
NVLink and summed on the device.  The Fr NTT at 2^22 is measured in the same run and reported in
the "ntt" object.  Inputs are resident in HBM for `value`; `e2e` goes through the public host-buffer
entry points with the host<->device copies inside the timed region.

The reference's own CPU implementation of this path (blst via midnight-curves) cannot be built
here (no Rust toolchain, crates not vendored), so both `cpu_baseline` and `--impl reference` time the
oracle's C restatement ("kind": "port") on the box's host cores -- a reported baseline, not a target.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BASE_SEED = 0xB200
SCALAR_SEED = 1
NTT_SEED = 2
LMAC_PER_POINT = 48000       # 16 windows x 10 Fp mul x 300 limb-MACs (SURVEY.md 8d accounting)
LMAC_PER_FR_MUL = 136
NTT_BYTES_PER_ELEM = 128     # 2 passes x (32 B read + 32 B write)


# ----------------------------------------------------------------------------------------------
# synthetic scalars (same definition as the checker's generator: limbs splitmix64(seed + 4i + j),
# top limb masked to 63 bits, one conditional subtraction of r)
# ----------------------------------------------------------------------------------------------
def synth_scalars_np(seed: int, start: int, n: int):
    import numpy as np

    x = (np.arange(4 * start, 4 * (start + n), dtype=np.uint64) + np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    z = x.copy()
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    k = z.reshape(n, 4)
    k[:, 3] &= np.uint64(0x7FFFFFFFFFFFFFFF)
    r = [np.uint64((R_MOD >> (64 * i)) & 0xFFFFFFFFFFFFFFFF) for i in range(4)]
    ge = np.zeros(n, dtype=bool)
    decided = np.zeros(n, dtype=bool)
    for i in (3, 2, 1, 0):
        gt, lt = k[:, i] > r[i], k[:, i] < r[i]
        ge |= gt & ~decided
        decided |= gt | lt
    ge |= ~decided                                # equal to r
    borrow = np.zeros(n, dtype=np.uint64)
    for i in range(4):
        sub = r[i] + borrow                       # r limbs are < 2^64 - 1, no wrap
        nb = (k[:, i] < sub).astype(np.uint64)
        k[:, i] = np.where(ge, k[:, i] - sub, k[:, i])
        borrow = nb
    return k


class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(fname, n_kernels):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the
    same command (profiles/), summed over the launches that make up one step of that kernel; None if absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", fname)))
        ks = d["kernels"][:n_kernels]
        return sum(k["traffic_bytes"] for k in ks) if len(ks) == n_kernels else None
    except Exception:
        return None


def load_oracle():
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    L = C.CDLL(so)
    L.orc_init()
    u64 = C.c_uint64
    L.orc_g1_msm.argtypes = [C.c_void_p, C.c_void_p, u64, C.c_void_p, C.c_int]
    L.orc_g1_synth_bases.argtypes = [u64, u64, u64, C.c_void_p, C.c_int]
    L.orc_ntt.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_max_threads.restype = C.c_int
    return L


def cpu_msm_sample(L, log_sample: int, reps: int = 1):
    """Times the CPU restatement on the first 2^log_sample points of the workload (all host threads)."""
    import numpy as np

    n = 1 << log_sample
    bases = np.empty(96 * n, dtype=np.uint8)
    L.orc_g1_synth_bases(BASE_SEED, 0, n, bases.ctypes.data, 0)
    sc = synth_scalars_np(SCALAR_SEED, 0, n)
    out = np.zeros(96, dtype=np.uint8)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        L.orc_g1_msm(bases.ctypes.data, sc.ctypes.data, n, out.ctypes.data, 0)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, best, bytes(out), bases, sc


def run_reference(args):
    """CPU arm: the oracle port on all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = load_oracle()
    cores = L.orc_max_threads()
    log_sample = min(args.cpu_sample_log_n, args.log_n)
    import numpy as np

    n = 1 << log_sample
    bases = np.empty(96 * n, dtype=np.uint8)
    L.orc_g1_synth_bases(BASE_SEED, 0, n, bases.ctypes.data, 0)
    sc = synth_scalars_np(SCALAR_SEED, 0, n)
    out = np.zeros(96, dtype=np.uint8)
    for _ in range(args.warmup):
        L.orc_g1_msm(bases.ctypes.data, sc.ctypes.data, n, out.ctypes.data, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        L.orc_g1_msm(bases.ctypes.data, sc.ctypes.data, n, out.ctypes.data, 0)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "first 2^%d of the 2^%d points per step (CPU restatement of Pippenger, not blst)" % (log_sample, args.log_n)
    line = {
        "impl": "reference", "metric": "g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64x6 Montgomery (integer)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    return {"workload": "BLS12-381 G1 MSM, 2^%d points, uniform 255-bit scalars (distribution U)" % args.log_n,
            "log_n": args.log_n, "bases": "a_i*G, a_i = splitmix64(0xB200 + i)", "scalars": "splitmix64 seed 1, reduced mod r",
            "sharding": "point-range over %d GPU(s), partial points all-gathered and summed on device" % n_gpus,
            "l2": "inputs (%.0f MiB bases + %.0f MiB scalars per GPU) exceed the 126 MB L2" %
                  (96.0 * (1 << args.log_n) / n_gpus / 2**20, 32.0 * (1 << args.log_n) / n_gpus / 2**20),
            "ntt": "Fr NTT 2^%d forward, natural order" % args.ntt_log_n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--ntt-log-n", type=int, default=22)
    ap.add_argument("--cpu-sample-log-n", type=int, default=19)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--acc-variant", type=int, default=-1, help="accumulate-kernel code variant (experiments)")
    ap.add_argument("--no-tables", action="store_true", help="do not build window tables at registration")
    ap.add_argument("--smax", type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    from importlib import import_module
    zdist = import_module("plutus-halo2-verifier-gen_b200.dist")
    zk.init(local_rank)
    lib = zk.lib()
    if args.window_bits or args.acc_variant >= 0 or args.smax or args.no_tables:
        lib.b200zk_set_msm_tuning(args.window_bits | ((args.acc_variant + 1) << 8 if args.acc_variant >= 0 else 0)
                                  | (0x8000 if args.no_tables else 0), args.smax)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- workload (untimed set-up)
    n_total = 1 << args.log_n
    start, end = zdist.shard_range(n_total, rank, world)
    n_local = end - start
    d_bases = torch.empty(n_local * 96, dtype=torch.uint8, device="cuda")
    zk.capi.check(lib.b200zk_g1_synth_bases_dev(BASE_SEED, start, n_local, d_bases.data_ptr(), stream))
    torch.cuda.synchronize()
    h = C.c_uint64(0)
    zk.capi.check(lib.b200zk_bases_register_dev(d_bases.data_ptr(), n_local, zk.FMT_MONT, 96, C.byref(h)))
    del d_bases
    torch.cuda.empty_cache()
    h_sc = torch.from_numpy(synth_scalars_np(SCALAR_SEED, start, n_local).view(np.uint8).reshape(-1)).pin_memory()
    d_sc = h_sc.cuda()
    h_out = torch.zeros(96, dtype=torch.uint8).pin_memory()
    msm = zdist.ShardedMSM(h.value, n_local, rank, world)
    zk.capi.set_profiling(True)

    # ---------------------------------------------------------------- device-resident timing
    for _ in range(args.warmup):
        msm.run_device(d_sc)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = zk.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        msm.run_device(d_sc)
    ev1.record()
    barrier()
    gpu_launches = zk.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    value = n_total / (ms_per_step * 1e-3)
    prof = zk.capi.get_profile()                      # phases of the last timed step on this rank
    acc_ms = max_over_ranks(prof.get("accumulate", 0.0))
    result_dev = bytes(msm.d_out.cpu().numpy())

    # ---------------------------------------------------------------- end to end (host buffers)
    for _ in range(2):
        msm.run_host(h_sc, h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        msm.run_host(h_sc, h_out)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    result_e2e = bytes(h_out.numpy())
    if world == 1:                                     # N = 1: also through the plain C-ABI host entry point
        out_c = C.create_string_buffer(96)
        zk.capi.check(lib.b200zk_msm_g1(h.value, 0, h_sc.data_ptr(), n_local, 0, zk.capi.addr(out_c)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            zk.capi.check(lib.b200zk_msm_g1(h.value, 0, h_sc.data_ptr(), n_local, 0, zk.capi.addr(out_c)))
        e2e_s = (time.perf_counter() - t0) / args.steps
        result_e2e = out_c.raw
    assert result_e2e == result_dev, "host-buffer path and device-resident path disagree"

    # ---------------------------------------------------------------- integer-pipe peak, measured in this run
    ops, ms = C.c_double(), C.c_double()
    zk.capi.check(lib.b200zk_microbench(7, 4000, C.byref(ops), C.byref(ms)))
    imad_peak = max_over_ranks(ops.value)             # IMAD.WIDE.U32 limb-MACs per second on one GPU
    achieved = n_local * LMAC_PER_POINT / (acc_ms * 1e-3) if acc_ms else 0.0

    # ---------------------------------------------------------------- NTT 2^22 (rank 0's GPU; replicas only)
    ntt = None
    if not args.no_ntt and rank == 0:
        ntt = bench_ntt(zk, lib, torch, np, args, stream, imad_peak)
    barrier()

    # ---------------------------------------------------------------- CPU baseline + parity on the sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        L = load_oracle()
        log_s = min(args.cpu_sample_log_n, args.log_n)
        pts_per_s, secs, cpu_pt, _, sc_s = cpu_msm_sample(L, log_s, reps=3)
        out_c = C.create_string_buffer(96)
        zk.capi.check(lib.b200zk_msm_g1(h.value, 0, sc_s.ctypes.data, 1 << log_s, 0, zk.capi.addr(out_c)))
        cpu = {"value": pts_per_s, "unit": "points/s", "cores": L.orc_max_threads(), "kind": "port",
               "sample": "first 2^%d of the 2^%d points, best of 3 runs (%.1f s each, all host threads); CPU restatement of "
                         "signed-window Pippenger, not blst" % (log_s, args.log_n, secs),
               "parity_on_sample": out_c.raw == cpu_pt}
        assert cpu["parity_on_sample"], "GPU MSM differs from the CPU oracle on the sample"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {
            "metric": "g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x12 (381-bit Montgomery, integer)", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": n_total / e2e_s, "unit": "points/s", "h2d_bytes_per_step": 32 * n_total,
                    "d2h_bytes_per_step": 96, "ms_per_step": e2e_s * 1e3,
                    "timed": "host clock around synchronous public calls (H2D scalars, MSM, exchange, D2H result)"},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "imad", "kernel": "msm_accumulate_kernel_v6", "achieved": achieved / 1e12,
                         "peak": imad_peak / 1e12, "unit": "Tlimb-MAC/s", "frac": achieved / imad_peak if imad_peak else None,
                         "traffic": ncu_traffic("r01_ncu_msm_accumulate_2p24.json", 1) if (args.log_n == 24 and world == 1) else None,
                         "traffic_note": "bytes per launch of the accumulate kernel (ncu capture of this command, profiles/); "
                                         "algorithmic bytes = (96 B base + 4 B entry) x points x rows = %.1f GB"
                                         % (100.0 * n_local * (prof.get("windows") or 0) / 1e9),
                         "note": "algorithmic work = %d limb-MACs/point x %d points per launch / accumulate-kernel time "
                                 "(CUDA events on the launching stream, last timed step, max over ranks); peak = IMAD.WIDE.U32 "
                                 "micro-benchmark measured in this run; window bits actually used: %s"
                                 % (LMAC_PER_POINT, n_local, prof.get("window_bits")),
                         "issued_imad_wide_frac": (2611.0 * n_local * (prof.get("windows") or 0) / (acc_ms * 1e-3)) / imad_peak
                         if (acc_ms and imad_peak) else None,
                         "issued_note": "IMAD.WIDE actually issued by the kernel (6 products x 289 + 2 squarings x 222 + one fused pair of products "
                                        "x 433 per mixed addition, one addition per point and table row) / time / peak: the pipe utilisation; "
                                        "`frac` above uses the fixed 16-window accounting of SURVEY 8d and exceeds it when window "
                                        "tables allow fewer rows",
                         "phases_ms": {k: prof.get(k) for k in ("sort", "accumulate", "tail")},
                         "hbm_crosscheck_gbs": n_local * 128 / (acc_ms * 1e-3) / 1e9 if acc_ms else None,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "cpu_baseline": cpu,
            "ntt": ntt,
            "result_compressed": zk.host.g1_compress(result_dev).hex(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_ntt(zk, lib, torch, np, args, stream, imad_peak):
    log_n = args.ntt_log_n
    n = 1 << log_n
    omega = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - log_n), R_MOD).to_bytes(32, "little")
    h_data = torch.from_numpy(synth_scalars_np(NTT_SEED, 0, n).view(np.uint8).reshape(-1)).pin_memory()
    d_data = h_data.cuda()
    for _ in range(args.warmup):
        zk.capi.check(lib.b200zk_ntt_fr_dev(d_data.data_ptr(), 1, log_n, zk.capi.addr(omega), 0, 0, stream))
    torch.cuda.synchronize()
    launches0 = zk.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        zk.capi.check(lib.b200zk_ntt_fr_dev(d_data.data_ptr(), 1, log_n, zk.capi.addr(omega), 0, 0, stream))
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = zk.launch_count() - launches0
    prof = zk.capi.get_profile()
    # end to end through the host-buffer entry point (in place on pinned memory)
    zk.capi.check(lib.b200zk_ntt_fr(h_data.data_ptr(), log_n, zk.capi.addr(omega), 0, 0))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        zk.capi.check(lib.b200zk_ntt_fr(h_data.data_ptr(), log_n, zk.capi.addr(omega), 0, 0))
    e2e_s = (time.perf_counter() - t0) / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    gbs = NTT_BYTES_PER_ELEM * n / (ms * 1e-3) / 1e9
    lmacs = (n // 2) * log_n * LMAC_PER_FR_MUL
    t_hbm = NTT_BYTES_PER_ELEM * n / (hbm_peak * 1e9)
    t_imad = lmacs / imad_peak if imad_peak else 0.0
    t_bound = max(t_hbm, t_imad)
    cpu = None
    if not args.no_cpu:
        L = load_oracle()
        buf = synth_scalars_np(NTT_SEED, 0, n)
        t0 = time.perf_counter()
        L.orc_ntt(buf.ctypes.data, log_n, omega, 0, None, None, 0)
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "elements/s", "cores": L.orc_max_threads(), "kind": "port",
               "sample": "one full 2^%d transform (%.2f s), CPU restatement of radix-2 NTT" % (log_n, dt)}
    return {
        "metric": "fr_ntt_elements_per_s", "log_n": log_n, "value": n / (ms * 1e-3), "unit": "elements/s", "ms_per_step": ms,
        "gpu_launches": launches, "passes_ms": prof.get("passes"),
        "e2e": {"value": n / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n,
                "ms_per_step": e2e_s * 1e3},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                     "traffic": ncu_traffic("r01_ncu_ntt_pass_2p22.json", 2) if log_n == 22 else None,
                     "imad": {"achieved_tlmac_s": lmacs / (ms * 1e-3) / 1e12, "peak_tlmac_s": imad_peak / 1e12,
                              "frac": (lmacs / (ms * 1e-3)) / imad_peak if imad_peak else None},
                     "binding": "imad" if t_imad > t_hbm else "hbm", "frac_of_binding_bound": t_bound / (ms * 1e-3),
                     "note": "algorithmic bytes = 128 B/element (2 passes x 32 B read + 32 B write); integer work = "
                             "(n/2) log2 n butterflies x 136 limb-MACs"},
        "cpu_baseline": cpu,
    }


if __name__ == "__main__":
    main()
