"""Replays the MSM + NTT call trace of one create_proof through the C ABI for the reference's
example circuits (shapes from SURVEY.md section 3.2 / 8d: commitments per proof from
/root/reference/docs/verifier_math.js:84-93 and the per-example counts at
examples/simple_mul.rs:91-92, atms.rs:85-86, atms_with_lookups.rs:110-111, ivc.rs:243-244).

The reference's prover cannot run here (no Rust), so this is the replay SURVEY 8d prescribes for the
"create_proof ms" metric: every commitment becomes one MSM of n = 2^k scalars against a resident
table (advice-like columns use the prover-like distribution S, quotient pieces / f / pi uniform),
every committed column costs one inverse NTT of size n and one coset NTT of size 2^(k+2), plus
one inverse coset NTT for the quotient.  CPU column: the checker's C port on all host threads
(one MSM and one NTT of each size timed, scaled by the call counts).

usage: python tools/circuit_bench.py [--circuits simple_mul,atms17,...] [--no-cpu]
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (scalar generator, oracle loader)
sys.path.insert(0, os.path.join(ROOT, "tools"))

#            name: (k, proof commitments, of which uniform (h pieces + f + pi + random poly))
CIRCUITS = {
    "simple_mul": (5, 10, 5),
    "lookup_table": (11, 20, 6),
    "atms14": (14, 18, 6),
    "atms17": (17, 18, 6),
    "atms19": (19, 18, 6),
    "atms_lookups17": (17, 22, 6),
    "sha256_19": (19, 26, 7),
    "ivc19": (19, 43, 7),
}


def prover_like(np, seed, n):
    """distribution S: 70% in {0,1}, 20% < 2^16, 10% uniform; the last 6 rows are blinding (uniform)"""
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    k = bench.synth_scalars_np(seed, 0, n)
    small = rng.integers(0, 1 << 16, n, dtype=np.uint64)
    bit = rng.integers(0, 2, n, dtype=np.uint64)
    m1, m2 = u < 0.7, (u >= 0.7) & (u < 0.9)
    m1[-6:] = False
    m2[-6:] = False
    k[m1, 0] = bit[m1]
    k[m2, 0] = small[m2]
    k[m1 | m2, 1:] = 0
    return k


def run_circuit(zk, lib, L, name, reps=3, check_all=False, threads=0):
    """One proof-shaped replay (host-buffer route and resident route) of circuit `name`; returns the record.  L = the CPU
    checker (None: no CPU column / parity).  check_all: every commitment is compared with the checker's MSM, not only the first."""
    import numpy as np
    import torch

    st = torch.cuda.current_stream().cuda_stream
    args = type("A", (), {"reps": reps})()
    k, ncom, nuni = CIRCUITS[name]
    n = 1 << k
    ek = k + 2
    # SRS table (synthetic, generated on the device) -> resident handle with window tables
    d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
    zk.capi.check(lib.b200zk_g1_synth_bases_dev(0xB200, 0, n, d_b.data_ptr(), st))
    torch.cuda.synchronize()
    h = C.c_uint64(0)
    zk.capi.check(lib.b200zk_bases_register_dev(d_b.data_ptr(), n, zk.FMT_MONT, 96, C.byref(h)))
    cols = [prover_like(np, 100 + i, n) for i in range(ncom - nuni)] + \
           [bench.synth_scalars_np(200 + i, 0, n) for i in range(nuni)]
    sc = torch.from_numpy(np.concatenate(cols).view(np.uint8).reshape(-1)).pin_memory()
    out = torch.zeros(96 * ncom, dtype=torch.uint8).pin_memory()
    omega_inv = pow(pow(zk.host.ROOT_OF_UNITY, 1 << (32 - k), zk.host.R_MOD), zk.host.R_MOD - 2, zk.host.R_MOD).to_bytes(32, "little")
    omega_ext = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - ek), zk.host.R_MOD)
    omega_ext_b = omega_ext.to_bytes(32, "little")
    omega_ext_inv = pow(omega_ext, zk.host.R_MOD - 2, zk.host.R_MOD).to_bytes(32, "little")
    g = zk.host.ZETA.to_bytes(32, "little")
    gi = pow(zk.host.ZETA, zk.host.R_MOD - 2, zk.host.R_MOD).to_bytes(32, "little")
    ncols = ncom - 3                                   # columns that go through the domain transforms
    lag = torch.from_numpy(np.concatenate(cols[:ncols]).view(np.uint8).reshape(-1)).pin_memory()
    ext = torch.zeros(32 * (1 << ek) * ncols, dtype=torch.uint8).pin_memory()

    def gpu_trace():
        # all commitments of the proof: one batched MSM call (host buffers in, 96-byte points out)
        zk.capi.check(lib.b200zk_msm_g1_batch(h.value, 0, sc.data_ptr(), n, ncom, 0, out.data_ptr()))
        # lagrange_to_coeff for every column, then coeff_to_extended, then one extended_to_coeff
        zk.capi.check(lib.b200zk_ntt_fr_batch(lag.data_ptr(), ncols, k, zk.capi.addr(omega_inv), zk.NTT_INVERSE_SCALE, 0))
        zk.capi.check(lib.b200zk_ntt_fr_batch(ext.data_ptr(), ncols, ek, zk.capi.addr(omega_ext_b), zk.NTT_COSET_IN, zk.capi.addr(g)))
        zk.capi.check(lib.b200zk_ntt_fr_batch(ext.data_ptr(), 1, ek, zk.capi.addr(omega_ext_inv),
                                              zk.NTT_INVERSE_SCALE | zk.NTT_COSET_OUT, zk.capi.addr(gi)))

    gpu_trace()
    l0 = zk.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        gpu_trace()
    gpu_ms = (time.perf_counter() - t0) / args.reps * 1e3
    launches = (zk.launch_count() - l0) // args.reps
    # MSM-only and NTT-only split
    t0 = time.perf_counter()
    for _ in range(args.reps):
        zk.capi.check(lib.b200zk_msm_g1_batch(h.value, 0, sc.data_ptr(), n, ncom, 0, out.data_ptr()))
    msm_ms = (time.perf_counter() - t0) / args.reps * 1e3
    rec = {"circuit": name, "k": k, "commitments": ncom, "ntt_columns": ncols, "gpu_trace_ms": gpu_ms, "gpu_msm_ms": msm_ms,
           "gpu_ntt_ms": gpu_ms - msm_ms, "gpu_launches": launches, "timed": "host clock, host buffers in and out"}
    # ---- the same proof with columns resident in HBM: witness columns go up once (Lagrange values),
    # everything between them and the commitments stays on the device: commit_lagrange, lagrange_to_coeff,
    # zero-padded coeff_to_extended, a gate program for the quotient numerator with 1/(X^n - 1) fused,
    # extended_to_coeff, commitment of the quotient pieces; only the G1 points come back.
    import random as _random
    from poly_bench import atms_like_program
    n_ext = 1 << ek
    rng = _random.Random(k)
    words, _muls, n_instr = atms_like_program(ncols, 4, max(4, ncols), rng)
    t_inv = [rng.randrange(1, zk.host.R_MOD) for _ in range(4)]
    gp = zk.host.GateProgram(words, [rng.randrange(zk.host.R_MOD), 0], [0, 1, -1, 2], ncols, k, ek, t_inv)
    d_cols = torch.empty(32 * n * ncom, dtype=torch.uint8, device="cuda")
    d_ext = torch.empty(32 * n_ext * ncols, dtype=torch.uint8, device="cuda")
    d_h = torch.empty(32 * n_ext, dtype=torch.uint8, device="cuda")
    d_pts = torch.empty(96 * (ncom + 3), dtype=torch.uint8, device="cuda")
    h_pts = torch.zeros(96 * (ncom + 3), dtype=torch.uint8).pin_memory()
    col_ptrs = (C.c_void_p * ncols)(*[d_ext.data_ptr() + 32 * n_ext * i for i in range(ncols)])
    MONT = zk.NTT_MONT

    side = torch.cuda.Stream()
    ev_rest = torch.cuda.Event()
    gA = max(1, ncom // 3) if 32 * n * ncom >= (256 << 20) else ncom   # below ~256 MiB one group: a second batch costs a second tail (k = 17: 15.3 -> 17.0 ms when split)
    phase_names = ("h2d_first_group", "to_montgomery_first_group", "commit_columns_with_rest_of_h2d", "lagrange_to_coeff", "zero_pad", "coeff_to_extended", "gate_program",
                   "extended_to_coeff", "commit_quotient", "d2h")

    def resident_trace(marks=None):
        def mark():
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)
        mark()
        # H2D once: ncom * n * 32 bytes.  The first third of the columns goes up on the main stream and is committed while the
        # rest crosses PCIe on a side stream (a prover hands its witness columns over as they are ready).
        bA = 32 * n * gA
        side.wait_stream(torch.cuda.current_stream())                                  # the previous pass still reads d_cols
        d_cols[:bA].copy_(sc[:bA], non_blocking=True)
        with torch.cuda.stream(side):
            d_cols[bA:].copy_(sc[bA:], non_blocking=True)
            ev_rest.record(side)
        mark()
        zk.capi.check(lib.b200zk_fr_convert_dev(d_cols.data_ptr(), d_cols.data_ptr(), n * gA, 1, st))
        mark()
        zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_cols.data_ptr(), n, gA, zk.FMT_MONT, 0, d_pts.data_ptr(), st))
        torch.cuda.current_stream().wait_event(ev_rest)
        if gA < ncom:
            zk.capi.check(lib.b200zk_fr_convert_dev(d_cols.data_ptr() + bA, d_cols.data_ptr() + bA, n * (ncom - gA), 1, st))
            zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_cols.data_ptr() + bA, n, ncom - gA, zk.FMT_MONT, 0, d_pts.data_ptr() + 96 * gA, st))
        mark()
        zk.capi.check(lib.b200zk_ntt_fr_dev(d_cols.data_ptr(), ncols, k, zk.capi.addr(omega_inv), zk.NTT_INVERSE_SCALE | MONT, 0, st))
        mark()
        zk.capi.check(lib.b200zk_fr_extend_dev(d_cols.data_ptr(), n, d_ext.data_ptr(), n_ext, ncols, st))
        mark()
        zk.capi.check(lib.b200zk_ntt_fr_dev(d_ext.data_ptr(), ncols, ek, zk.capi.addr(omega_ext_b), zk.NTT_COSET_IN | MONT, zk.capi.addr(g), st))
        mark()
        zk.capi.check(lib.b200zk_gate_program_run_dev(gp.handle, C.addressof(col_ptrs), d_h.data_ptr(), 0, st))
        mark()
        zk.capi.check(lib.b200zk_ntt_fr_dev(d_h.data_ptr(), 1, ek, zk.capi.addr(omega_ext_inv),
                                            zk.NTT_INVERSE_SCALE | zk.NTT_COSET_OUT | MONT, zk.capi.addr(gi), st))
        mark()
        zk.capi.check(lib.b200zk_msm_g1_dev(h.value, 0, d_h.data_ptr(), n, 3, zk.FMT_MONT, 0, d_pts.data_ptr() + 96 * ncom, st))
        mark()
        h_pts.copy_(d_pts, non_blocking=True)
        mark()
        torch.cuda.synchronize()

    resident_trace()
    assert bytes(h_pts[:96 * ncom].numpy()) == bytes(out.numpy()), "resident commitments differ from the host-buffer path"
    t0 = time.perf_counter()
    for _ in range(args.reps):
        resident_trace()
    rec["gpu_resident_trace_ms"] = (time.perf_counter() - t0) / args.reps * 1e3
    marks = []
    resident_trace(marks)                                                              # one more pass with an event after every step
    rec["resident_phases_ms"] = {nm: marks[i].elapsed_time(marks[i + 1]) for i, nm in enumerate(phase_names)}
    rec["resident_note"] = ("columns uploaded once and kept in HBM: %d commitments + %d inverse NTTs + %d coset NTTs of 2^%d + gate program "
                            "(%d instructions over %d columns) + quotient inverse NTT + 3 quotient commitments; D2H = %d bytes"
                            % (ncom, ncols, ncols, ek, n_instr, ncols, 96 * (ncom + 3)))
    # ---- the opening argument over the same resident coefficient columns (KZGCommitmentScheme::multi_open, the last prover
    # phase): every column at x, a third of them also at omega x, a sixth at omega^-1 x -> three point sets, two more commitments
    w_k = pow(zk.host.ROOT_OF_UNITY, 1 << (32 - k), zk.host.R_MOD)
    x_pt = rng.randrange(zk.host.R_MOD)
    rots = [x_pt, x_pt * w_k % zk.host.R_MOD, x_pt * pow(w_k, zk.host.R_MOD - 2, zk.host.R_MOD) % zk.host.R_MOD]
    queries = [(i, rots[0]) for i in range(ncols)] + [(i, rots[1]) for i in range(0, ncols, 3)] + [(i, rots[2]) for i in range(1, ncols, 6)]
    poly_ptrs = (C.c_void_p * ncols)(*[d_cols.data_ptr() + 32 * n * i for i in range(ncols)])
    q_poly = (C.c_uint32 * len(queries))(*[q[0] for q in queries])
    q_pts = b"".join(zk.host.fr_bytes(q[1]) for q in queries)
    proof_buf = C.create_string_buffer(96 + 32 * len(queries))
    proof_len = C.c_size_t(0)

    def open_step():
        tr = zk.host.Transcript()
        tr.common_scalar(k)
        zk.capi.check(lib.b200zk_h2mo_open_dev(h.value, tr.handle, C.addressof(poly_ptrs), ncols, n, C.addressof(q_poly), zk.capi.addr(q_pts),
                                               len(queries), zk.capi.addr(proof_buf), len(proof_buf), C.byref(proof_len)))
        tr.free()
        return proof_buf.raw[:proof_len.value]

    first = open_step()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        again = open_step()
    rec["gpu_multi_open_ms"] = (time.perf_counter() - t0) / args.reps * 1e3
    assert again == first and len(first) == 96 + 32 * 3, "the opening proof is not reproducible"
    rec["gpu_resident_proof_ms"] = rec["gpu_resident_trace_ms"] + rec["gpu_multi_open_ms"]
    rec["multi_open_note"] = ("b200zk_h2mo_open_dev over the %d resident coefficient columns, %d queries in 3 point sets: linear combinations, "
                              "Kate divisions, evaluations and the two commitments (f, pi) on the GPU; transcript and set construction "
                              "on the host; 192 proof bytes come back" % (ncols, len(queries)))
    gp.release()
    del d_cols, d_ext, d_h
    if L is not None:
        bases = np.empty(96 * n, dtype=np.uint8)
        L.orc_g1_synth_bases(0xB200, 0, n, bases.ctypes.data, threads)
        o = np.zeros(96, dtype=np.uint8)
        tt = []
        for col in (cols[0], cols[-1]):               # one prover-like and one uniform commitment
            t0 = time.perf_counter()
            L.orc_g1_msm(bases.ctypes.data, col.ctypes.data, n, o.ctypes.data, threads)
            tt.append(time.perf_counter() - t0)
        checked = range(ncom) if check_all else (0,)
        good = True
        for j in checked:
            L.orc_g1_msm(bases.ctypes.data, cols[j].ctypes.data, n, o.ctypes.data, threads)
            good = good and bytes(out[96 * j:96 * j + 96].numpy()) == bytes(o)
        rec["parity_commitments_checked"] = len(checked)
        rec["parity"] = good
        assert good, "a commitment of the %s trace differs from the CPU checker" % name
        cpu_msm = tt[0] * (ncom - nuni) + tt[1] * nuni
        buf = cols[-1].copy()
        t0 = time.perf_counter()
        L.orc_ntt(buf.ctypes.data, k, omega_inv, 1, None, None, threads)
        t_small = time.perf_counter() - t0
        big = np.zeros((1 << ek, 4), dtype=np.uint64)
        big[:n] = cols[-1]
        t0 = time.perf_counter()
        L.orc_ntt(big.ctypes.data, ek, omega_ext_b, 0, g, None, threads)
        t_big = time.perf_counter() - t0
        cpu_ntt = t_small * ncols + t_big * (ncols + 1)
        rec.update({"cpu_trace_ms": (cpu_msm + cpu_ntt) * 1e3, "cpu_msm_ms": cpu_msm * 1e3, "cpu_ntt_ms": cpu_ntt * 1e3,
                    "cpu_cores": threads or L.orc_max_threads(), "cpu_kind": "port (checker's C restatement, sampled: 2 MSMs + 2 NTTs scaled by call counts)"})
    zk.capi.check(lib.b200zk_bases_release(h.value))
    del d_b
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--circuits", default=",".join(CIRCUITS))
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--check-all", action="store_true", help="compare every commitment with the CPU checker")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zk.init(0)
    lib = zk.lib()
    L = None if args.no_cpu else bench.load_oracle()
    for name in args.circuits.split(","):
        rec = run_circuit(zk, lib, L, name, args.reps, args.check_all, bench.host_threads())
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
