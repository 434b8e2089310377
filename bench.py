#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 MSM / NTT backend.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the CPU path, timed on the host cores

Metric (BASELINE.json): BLS12-381 G1 MSM points/s on 2^24 synthetic points with uniform scalars
(distribution "U"); a step is one full MSM.  With N > 1 GPUs the same 2^24
points are sharded by point range (strong scaling) and the N partial sums meet in GPU 0's HBM
through peer stores issued by the MSM's last kernel (the last arriver folds them).  The Fr NTT at 2^22 is measured in the same run and reported in
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
    L.orc_fr_dot_u64.argtypes = [C.c_void_p, C.c_void_p, u64, C.c_void_p]
    L.orc_g1_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_g1_generator.argtypes = [C.c_void_p]
    return L


def host_threads() -> int:
    """Every host thread the box has: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would silently turn the CPU arm
    into a single-core run, so the thread count is passed to the oracle explicitly."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def dlog_identity_point(L, scalars_np, seed, start, n):
    """(sum_i s_i * a_i mod r) * G for the synthetic bases P_i = a_i * G, a_i = splitmix64(seed + start + i): what any correct MSM of
    the workload must return, computed without touching a single base point (one dot product + one scalar multiplication)."""
    import numpy as np

    x = (np.arange(start, start + n, dtype=np.uint64) + np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    z = x.copy()
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    a = np.ascontiguousarray(z ^ (z >> np.uint64(31)))
    dot = np.zeros(32, dtype=np.uint8)
    sc = np.ascontiguousarray(scalars_np)
    L.orc_fr_dot_u64(sc.ctypes.data, a.ctypes.data, n, dot.ctypes.data)
    gen = np.zeros(96, dtype=np.uint8)
    L.orc_g1_generator(gen.ctypes.data)
    out = np.zeros(96, dtype=np.uint8)
    L.orc_g1_mul(gen.ctypes.data, dot.ctypes.data, out.ctypes.data)
    return bytes(out)


def cpu_msm(L, log_n: int, threads: int, bases_ptr=None):
    """One MSM of the first 2^log_n points of the workload on the CPU checker.  bases_ptr: affine wire-format bases already
    in host memory (else they are synthesised on the CPU).  Returns (seconds, result bytes, scalars)."""
    import numpy as np

    n = 1 << log_n
    keep = None
    if bases_ptr is None:
        keep = np.empty(96 * n, dtype=np.uint8)
        L.orc_g1_synth_bases(BASE_SEED, 0, n, keep.ctypes.data, threads)
        bases_ptr = keep.ctypes.data
    sc = synth_scalars_np(SCALAR_SEED, 0, n)
    out = np.zeros(96, dtype=np.uint8)
    t0 = time.perf_counter()
    L.orc_g1_msm(bases_ptr, sc.ctypes.data, n, out.ctypes.data, threads)
    return time.perf_counter() - t0, bytes(out), sc


def run_reference(args):
    """CPU arm: the checker's C port of the path on every host thread.  Timed steps run the FULL workload when K of them fit the
    time budget (then same_config is true); otherwise each step is the largest power-of-two sample that fits, and the line
    says so.  Warm-up steps run a 2^20 sample (there is nothing to warm on a CPU; they are not timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    L = load_oracle()
    threads = host_threads()
    t_probe, _, _ = cpu_msm(L, min(20, args.log_n), threads)
    est_full = t_probe * (1 << max(0, args.log_n - 20)) * 0.85       # Pippenger: a little cheaper per point at larger n
    log_sample = args.log_n
    while log_sample > 16 and est_full * args.steps / (1 << (args.log_n - log_sample)) > args.cpu_budget_s:
        log_sample -= 1
    if args.cpu_sample_log_n:
        log_sample = min(args.cpu_sample_log_n, args.log_n)
    n = 1 << log_sample
    bases = np.empty(96 * n, dtype=np.uint8)
    t0 = time.perf_counter()
    L.orc_g1_synth_bases(BASE_SEED, 0, n, bases.ctypes.data, threads)
    synth_s = time.perf_counter() - t0
    sc = synth_scalars_np(SCALAR_SEED, 0, n)
    out = np.zeros(96, dtype=np.uint8)
    nw = 1 << min(20, log_sample)
    for _ in range(args.warmup):
        L.orc_g1_msm(bases.ctypes.data, sc.ctypes.data, nw, out.ctypes.data, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        L.orc_g1_msm(bases.ctypes.data, sc.ctypes.data, n, out.ctypes.data, threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    same = log_sample == args.log_n
    parity = dlog_identity_point(L, sc, BASE_SEED, 0, n) == bytes(out)
    sample = ("the full 2^%d-point workload per timed step" % args.log_n if same else
              "first 2^%d of the 2^%d points per timed step (K full-size steps exceed the %d s budget)" % (log_sample, args.log_n, args.cpu_budget_s)) + \
             "; %d host threads; CPU restatement of signed-window Pippenger, not blst; bases synthesised on the CPU in %.0f s (untimed)" % (threads, synth_s)
    line = {
        "impl": "reference", "metric": "g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64x6 Montgomery (integer)", "data": "synthetic",
        "config": workload_config(args, 1), "same_config": same,
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample,
                         "self_check_dlog_identity": parity},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    return {"workload": "BLS12-381 G1 MSM, 2^%d points, uniform 255-bit scalars (distribution U)" % args.log_n,
            "log_n": args.log_n, "bases": "a_i*G, a_i = splitmix64(0xB200 + i)", "scalars": "splitmix64 seed 1, reduced mod r",
            "sharding": "point-range over %d GPU(s); the per-GPU partial sums meet in GPU 0's HBM through peer stores issued by the "
                        "MSM's last kernel, the last arriver folds them" % n_gpus,
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
    ap.add_argument("--cpu-sample-log-n", type=int, default=0, help="force the size of the reference arm's timed steps (0 = full size if it fits)")
    ap.add_argument("--cpu-budget-s", type=int, default=420, help="time budget of the reference arm's timed steps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (and the full-size oracle parity check)")
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-circuits", action="store_true", help="skip the proof-shaped replays (BASELINE's create_proof metric)")
    ap.add_argument("--no-single-process", action="store_true", help="N > 1: skip the one-process-drives-all-GPUs measurement")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: how the partial sums meet")
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
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")      # host-side waits that must not keep a kernel spinning on the GPUs
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zdist = importlib.import_module("plutus-halo2-verifier-gen_b200.dist")
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
    free0 = torch.cuda.mem_get_info()[0]
    t0 = time.perf_counter()
    zk.capi.check(lib.b200zk_bases_register_dev(d_bases.data_ptr(), n_local, zk.FMT_MONT, 96, C.byref(h)))
    table_build_ms = (time.perf_counter() - t0) * 1e3
    table_bytes = free0 - torch.cuda.mem_get_info()[0]
    del d_bases
    torch.cuda.empty_cache()
    sc_np = synth_scalars_np(SCALAR_SEED, start, n_local)
    h_sc = torch.from_numpy(sc_np.view(np.uint8).reshape(-1)).pin_memory()
    d_sc = h_sc.cuda()
    h_out = torch.zeros(96, dtype=torch.uint8).pin_memory()
    msm = zdist.ShardedMSM(h.value, n_local, rank, world, exchange=args.exchange)
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
    sort_ms = max_over_ranks(prof.get("sort", 0.0))
    tail_ms = max_over_ranks(prof.get("tail", 0.0))
    result_dev = bytes(msm.d_out.cpu().numpy())

    # ---------------------------------------------------------------- end to end (host buffers, public C-ABI call)
    def e2e_call():
        if world == 1:
            zk.capi.check(lib.b200zk_msm_g1(h.value, 0, h_sc.data_ptr(), n_local, 0, h_out.data_ptr()))
        else:
            msm.run_host(h_sc, h_out)

    for _ in range(2):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    result_e2e = bytes(h_out.numpy())
    assert result_e2e == result_dev, "host-buffer path and device-resident path disagree"
    assert not msm.timed_out(), "a rank never delivered its partial sum"

    # ---------------------------------------------------------------- integer-pipe peak, measured in this run
    ops, ms = C.c_double(), C.c_double()
    zk.capi.check(lib.b200zk_microbench(7, 4000, C.byref(ops), C.byref(ms)))
    imad_peak = max_over_ranks(ops.value)             # IMAD.WIDE.U32 limb-MACs per second on one GPU
    rows = prof.get("windows") or 0
    IMAD_PER_MADD = 2611.0                            # 6 products x 289 + 2 squarings x 222 + one fused pair x 433
    issued = IMAD_PER_MADD * n_local * rows           # IMAD.WIDE the accumulate kernel issues per launch
    achieved = issued / (acc_ms * 1e-3) if acc_ms else 0.0
    fixed = n_local * LMAC_PER_POINT / (acc_ms * 1e-3) if acc_ms else 0.0

    # ---------------------------------------------------------------- NTT 2^22 (rank 0's GPU; replicas only)
    ntt = None
    if not args.no_ntt and rank == 0:
        ntt = bench_ntt(zk, lib, torch, np, args, stream, imad_peak)
    barrier()

    # ---------------------------------------------------------------- parity of the full-size result + CPU baseline
    parity, cpu = None, None
    if rank == 0:
        L = load_oracle()
        threads = host_threads()
        full_sc = sc_np if world == 1 else synth_scalars_np(SCALAR_SEED, 0, n_total)
        t0 = time.perf_counter()
        want = dlog_identity_point(L, full_sc, BASE_SEED, 0, n_total)
        parity = {"dlog_identity": want == result_dev, "dlog_identity_s": time.perf_counter() - t0,
                  "note": "MSM(s, a_i*G) == (sum s_i a_i mod r)*G over all 2^%d points, computed by the CPU checker without the bases" % args.log_n}
        assert parity["dlog_identity"], "GPU MSM violates the discrete-log identity on the full workload"
        if world == 1 and not args.no_cpu:
            # the full workload on the CPU, over the very bases the GPU used (read back from the resident table): the CPU baseline
            # and an independent full-size parity check in one run
            pts = np.empty(96 * n_total, dtype=np.uint8)
            zk.capi.check(lib.b200zk_bases_read(h.value, 0, n_total, pts.ctypes.data))
            secs, cpu_pt, _ = cpu_msm(L, args.log_n, threads, pts.ctypes.data)
            del pts
            parity["oracle_full_msm"] = cpu_pt == result_dev
            assert parity["oracle_full_msm"], "GPU MSM differs from the CPU oracle on the full workload"
            cpu = {"value": n_total / secs, "unit": "points/s", "cores": threads, "kind": "port",
                   "sample": "the full 2^%d-point workload, one run (%.1f s, %d host threads) over the bases read back from the GPU's table; CPU "
                             "restatement of signed-window Pippenger, not blst" % (args.log_n, secs, threads)}
        parity["parity_full"] = all(v for k, v in parity.items() if isinstance(v, bool))

    # ---------------------------------------------------------------- proof-shaped replays (BASELINE metric iii)
    circuits = None
    if rank == 0 and world == 1 and not args.no_circuits:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import circuit_bench
        L = None if args.no_cpu else load_oracle()
        circuits = {"note": "MSM + NTT call trace of one create_proof replayed through the C ABI (the Rust prover cannot run here): "
                            "host-buffer route and resident route; CPU column = the checker's port, sampled calls scaled by the call counts",
                    "traces": [circuit_bench.run_circuit(zk, lib, L, name, 3, True, host_threads()) for name in ("atms17", "atms19")]}

    # ---------------------------------------------------------------- batched verification up to the pairing (BASELINE configs[4])
    verify = None
    if rank == 0 and world == 1 and not args.no_circuits:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import verify_bench
        verify = verify_bench.run_batched_verify(zk, None if args.no_cpu else load_oracle())
        assert verify["accepted_s_left_eq_right"] in (True, None), "the batched guards do not verify"

    # ---------------------------------------------------------------- N > 1: ONE process driving all N GPUs through the plain C ABI
    single = None
    if world > 1 and not args.no_single_process:
        msm.close()
        zk.capi.check(lib.b200zk_bases_release(h.value))
        del d_sc
        torch.cuda.empty_cache()
        dist.barrier(group=cpu_group)
        if rank == 0:
            single = bench_single_process(zk, lib, torch, np, args, world, result_dev)
        dist.barrier(group=cpu_group)                   # the other ranks wait on the CPU, their GPUs are idle

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
                    "timed": "host clock around synchronous public calls (H2D scalars from pinned memory, MSM, exchange, D2H result)"},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "imad", "kernel": "msm_accumulate_kernel_v6", "achieved": achieved / 1e12,
                         "peak": imad_peak / 1e12, "unit": "T IMAD.WIDE/s", "frac": achieved / imad_peak if imad_peak else None,
                         "note": "achieved = IMAD.WIDE.U32 the kernel issues per launch (2611 per mixed addition x %d points x %d table rows) / "
                                 "its launch time (CUDA events on the launching stream, last timed step, max over ranks); peak = the "
                                 "IMAD.WIDE.U32 micro-benchmark run in this process (one limb-MAC per instruction)" % (n_local, rows),
                         "frac_whole_step": (issued / (ms_per_step * 1e-3)) / imad_peak if imad_peak else None,
                         "frac_fixed_accounting": fixed / imad_peak if imad_peak else None,
                         "fixed_accounting_note": "SURVEY 8d's size-independent accounting (48 000 limb-MACs/point = 16 windows x 10 products x 300) / "
                                                  "accumulate time / peak; exceeds 1 because window tables need only %d rows and the additions are "
                                                  "cheaper than 10 plain products -- a comparison figure, not a utilisation" % rows,
                         "traffic": ncu_traffic("r02_ncu_msm_acc.json", 1) if (args.log_n == 24 and world == 1) else None,
                         "traffic_source": "static_profile: profiles/r02_ncu_msm_acc.json (dram__bytes_read.sum + dram__bytes_write.sum of one launch, "
                                           "ncu --set full capture of `bench.py --steps 3 --warmup 3 --no-cpu --no-circuits`, tools/profile_round.sh), "
                                           "not measured in this run",
                         "algorithmic_bytes": 100.0 * n_local * rows,
                         "window_bits": prof.get("window_bits"), "table_rows": rows,
                         "phases_ms": {"sort": sort_ms, "accumulate": acc_ms, "tail": tail_ms},
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "table": {"build_ms": table_build_ms, "bytes": table_bytes,
                      "note": "window tables (rows 2^(c*w) * P_i) are built once per registered SRS table and are NOT inside any timed region; "
                              "a ParamsKZG holds two such tables (g, g_lagrange)"},
            "parity": parity,
            "cpu_baseline": cpu,
            "ntt": ntt,
            "circuits": circuits,
            "batched_verify": verify,
            "single_process": single,
            "result_compressed": zk.host.g1_compress(result_dev).hex(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_single_process(zk, lib, torch, np, args, n_gpus, expect):
    """The path a single-process Rust prover reaches: b200zk_init_devices + plain b200zk_msm_g1 against a table partitioned
    over all GPUs (worker thread per GPU, peer-store exchange).  Timed on the host clock around the synchronous calls."""
    zk.shutdown()
    zk.capi.init_devices(None, n_gpus)
    n = 1 << args.log_n
    torch.cuda.set_device(0)
    d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda:0")
    zk.capi.check(lib.b200zk_g1_synth_bases_dev(BASE_SEED, 0, n, d_b.data_ptr(), None))
    torch.cuda.synchronize()
    h = C.c_uint64(0)
    t0 = time.perf_counter()
    zk.capi.check(lib.b200zk_bases_register_dev(d_b.data_ptr(), n, zk.FMT_MONT | zk.capi.BASES_SHARD, 96, C.byref(h)))
    build_ms = (time.perf_counter() - t0) * 1e3
    del d_b
    torch.cuda.empty_cache()
    sc = synth_scalars_np(SCALAR_SEED, 0, n)
    h_sc = torch.from_numpy(sc.view(np.uint8).reshape(-1)).pin_memory()
    out = C.create_string_buffer(96)
    for _ in range(3):
        zk.capi.check(lib.b200zk_msm_g1(h.value, 0, h_sc.data_ptr(), n, 0, zk.capi.addr(out)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        zk.capi.check(lib.b200zk_msm_g1(h.value, 0, h_sc.data_ptr(), n, 0, zk.capi.addr(out)))
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    ok_host = out.raw == expect
    # resident: every GPU already holds its slice of the scalars
    layout, _ = zk.capi.bases_layout(h.value)
    slices = []
    for dev, s0, cnt in layout:
        t = torch.empty(32 * cnt, dtype=torch.uint8, device="cuda:%d" % dev)
        t.copy_(h_sc[32 * s0:32 * (s0 + cnt)])
        slices.append(t)
    for d in range(n_gpus):
        torch.cuda.synchronize(d)
    ptrs = (C.c_void_p * len(slices))(*[t.data_ptr() for t in slices])
    for _ in range(3):
        zk.capi.check(lib.b200zk_msm_g1_sharded_dev(h.value, C.addressof(ptrs), len(slices), 0, zk.capi.addr(out)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        zk.capi.check(lib.b200zk_msm_g1_sharded_dev(h.value, C.addressof(ptrs), len(slices), 0, zk.capi.addr(out)))
    res_ms = (time.perf_counter() - t0) / args.steps * 1e3
    ok_res = out.raw == expect
    zk.capi.check(lib.b200zk_bases_release(h.value))
    assert ok_host and ok_res, "single-process multi-GPU result differs from the per-rank path"
    del slices, h_sc
    ntt = bench_sharded_ntt(zk, lib, torch, np, args, n_gpus) if (n_gpus & (n_gpus - 1)) == 0 and not args.no_ntt else None
    # a proof's columns over the GPUs: the host-buffer route of the atms k = 19 trace (18 commitments as one pointer-less batch
    # call, 15 + 15 + 1 transforms as batch calls) with all N GPUs bound -- tables replicated, columns dealt out, no exchange
    proof = None
    if not args.no_circuits:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import circuit_bench
        rec = circuit_bench.run_circuit(zk, lib, None, "atms19", 3, False, 0)
        proof = {k: rec[k] for k in ("circuit", "k", "commitments", "ntt_columns", "gpu_trace_ms", "gpu_msm_ms", "gpu_ntt_ms")}
        proof["note"] = "host buffers in and out, host clock; compare with circuits.traces[atms19].gpu_trace_ms of the 1-GPU line"
    verify = None
    if not args.no_circuits:
        import verify_bench
        verify = verify_bench.run_batched_verify(zk, None if args.no_cpu else load_oracle())
    return {"n_gpus": n_gpus, "ntt_sharded": ntt, "proof_columns_over_gpus": proof, "batched_verify": verify, "e2e_single_process": {"value": n / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                                                     "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 96},
            "resident_single_process": {"value": n / (res_ms * 1e-3), "unit": "points/s", "ms_per_step": res_ms},
            "table_build_ms": build_ms, "parity_with_per_rank_path": True,
            "note": "one process, b200zk_init_devices(%d): b200zk_msm_g1 from pinned host scalars / b200zk_msm_g1_sharded_dev with resident "
                    "slices; host clock around the synchronous calls (includes the fan-out to the per-GPU worker threads)" % n_gpus}


def bench_sharded_ntt(zk, lib, torch, np, args, n_gpus, log_n=24):
    """ONE Fr NTT of 2^24 points over all GPUs of the single process (four-step, the exchange fused into the column pass as
    peer stores): resident form (blocks stay in HBM) and host-buffer form (both transpositions ride on the strided copies)."""
    n = 1 << log_n
    omega = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - log_n), R_MOD).to_bytes(32, "little")
    lr, lc = C.c_uint32(0), C.c_uint32(0)
    zk.capi.check(lib.b200zk_ntt_sharded_layout(log_n, C.byref(lr), C.byref(lc)))
    R, Cc = 1 << lr.value, 1 << lc.value
    Cl = Cc // n_gpus
    data = synth_scalars_np(NTT_SEED, 0, n)
    x = data.view(np.uint8).reshape(R, Cc, 32)
    ins = [torch.from_numpy(np.ascontiguousarray(x[:, g * Cl:(g + 1) * Cl, :]).reshape(-1)).to("cuda:%d" % g) for g in range(n_gpus)]
    outs = [torch.empty_like(t) for t in ins]
    pin = (C.c_void_p * n_gpus)(*[t.data_ptr() for t in ins])
    pout = (C.c_void_p * n_gpus)(*[t.data_ptr() for t in outs])
    for _ in range(3):
        zk.capi.check(lib.b200zk_ntt_fr_sharded_dev(C.addressof(pin), C.addressof(pout), n_gpus, log_n, zk.capi.addr(omega), 0, None))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        zk.capi.check(lib.b200zk_ntt_fr_sharded_dev(C.addressof(pin), C.addressof(pout), n_gpus, log_n, zk.capi.addr(omega), 0, None))
    res_ms = (time.perf_counter() - t0) / args.steps * 1e3
    # X[0] = sum of the inputs: one output any correct transform must produce (the full parity of this path is in the tests)
    y0 = outs[0].cpu().numpy()[:32].tobytes()                                    # block 0, [k2 = 0][k1 = 0] = X[0]
    acc = 0
    for j in range(4):
        lo = int((data[:, j] & np.uint64(0xFFFFFFFF)).sum(dtype=np.uint64))      # 2^24 x 2^32 < 2^64: no wrap
        hi = int((data[:, j] >> np.uint64(32)).sum(dtype=np.uint64))
        acc += (lo + (hi << 32)) << (64 * j)
    ok0 = int.from_bytes(y0, "little") == acc % R_MOD
    h_data = torch.from_numpy(data.view(np.uint8).reshape(-1).copy()).pin_memory()
    zk.capi.check(lib.b200zk_ntt_fr(h_data.data_ptr(), log_n, zk.capi.addr(omega), 0, None))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        zk.capi.check(lib.b200zk_ntt_fr(h_data.data_ptr(), log_n, zk.capi.addr(omega), 0, None))
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    assert ok0, "sharded NTT: X[0] is not the sum of the inputs"
    return {"log_n": log_n, "resident_ms": res_ms, "resident_elements_per_s": n / (res_ms * 1e-3), "e2e_ms": e2e_ms,
            "e2e_elements_per_s": n / (e2e_ms * 1e-3), "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n, "x0_check": ok0,
            "note": "host clock around the synchronous calls; resident: blocks [R][C/G] in, [C][R/G] out per GPU (R = 2^%d, C = 2^%d); "
                    "full parity of this path against the CPU oracle is in tests/multi/run_multi.py (2^13..2^23)" % (lr.value, lc.value)}


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
    issued = (n // 2) * log_n * 113          # IMAD.WIDE the kernels issue: one product per butterfly, 113 per Fr product
    t_hbm = NTT_BYTES_PER_ELEM * n / (hbm_peak * 1e9)
    t_imad = lmacs / imad_peak if imad_peak else 0.0
    t_bound = max(t_hbm, t_imad)
    cpu = None
    if not args.no_cpu:
        L = load_oracle()
        buf = synth_scalars_np(NTT_SEED, 0, n)
        t0 = time.perf_counter()
        L.orc_ntt(buf.ctypes.data, log_n, omega, 0, None, None, host_threads())
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "elements/s", "cores": host_threads(), "kind": "port",
               "sample": "one full 2^%d transform (%.2f s), CPU restatement of radix-2 NTT" % (log_n, dt)}
    return {
        "metric": "fr_ntt_elements_per_s", "log_n": log_n, "value": n / (ms * 1e-3), "unit": "elements/s", "ms_per_step": ms,
        "gpu_launches": launches, "passes_ms": prof.get("passes"),
        "e2e": {"value": n / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n,
                "ms_per_step": e2e_s * 1e3},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                     "traffic": ncu_traffic("r02_ncu_ntt.json", 2) if log_n == 22 else None,
                     "traffic_source": "static_profile: profiles/r02_ncu_ntt.json (both passes of one 2^22 transform), not measured in this run",
                     "imad": {"achieved_t_imad_wide_s": issued / (ms * 1e-3) / 1e12, "peak_t_imad_wide_s": imad_peak / 1e12,
                              "frac": (issued / (ms * 1e-3)) / imad_peak if imad_peak else None,
                              "frac_fixed_accounting": (lmacs / (ms * 1e-3)) / imad_peak if imad_peak else None,
                              "note": "frac = IMAD.WIDE.U32 actually issued ((n/2) log2 n products x 113) / time / measured peak; the fixed "
                                      "accounting of SURVEY 8d counts 136 limb-MACs per product"},
                     "binding": "imad" if t_imad > t_hbm else "hbm", "frac_of_binding_bound_fixed_accounting": t_bound / (ms * 1e-3),
                     "note": "algorithmic bytes = 128 B/element (2 passes x 32 B read + 32 B write); integer work = "
                             "(n/2) log2 n butterflies x 136 limb-MACs"},
        "cpu_baseline": cpu,
    }


if __name__ == "__main__":
    main()
