"""Measurement of the components next to the hot path on one GPU, device-resident, CUDA events on
the launching stream: Fr vector kernels (pointwise, linear combination, batched inversion, running product,
evaluation + Kate division), a circuit-shaped gate program over the extended domain, SRS generation, fixed-base
multiplication and batched decompression.  Each line carries the algorithmic HBM bytes and Fr products per
element, the achieved GB/s against MEASURED_PEAKS.json and the achieved limb-MACs/s against the IMAD.WIDE
micro-benchmark, and names the binding bound.
usage: python tools/poly_bench.py [--sizes 20,22,24] [--reps 5]"""
import argparse
import ctypes as C
import importlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

R = bench.R_MOD


def fr(x):
    return (x % R).to_bytes(32, "little")


def atms_like_program(n_cols, n_rot, n_gates, rng):
    """A quotient numerator shaped like the example circuits' gates: per gate a selector times a degree-3 product
    of advice cells plus linear terms, folded with the challenge y (const 0) by Horner steps."""
    prog = []
    K = lambda i: (0 << 28) | i
    Rg = lambda i: (1 << 28) | i
    Col = lambda c, r: (2 << 28) | (c << 12) | r
    prog.append((7, 0, K(1), 0, 0))                                   # acc = 0 (const 1 is zero)
    for g in range(n_gates):
        a, b, c, q = (rng.randrange(n_cols) for _ in range(4))
        prog.append((2, 1, Col(a, rng.randrange(n_rot)), Col(b, rng.randrange(n_rot)), 0))
        prog.append((6, 1, Rg(1), Col(c, rng.randrange(n_rot)), Col(a, 0)))
        prog.append((1, 1, Rg(1), Col(b, 0), 0))
        prog.append((2, 1, Rg(1), Col(q, 0), 0))
        prog.append((6, 0, Rg(0), K(0), Rg(1)))                        # acc = acc * y + gate
    words = []
    for (op, dst, x, y, z) in prog:
        words += [op | (dst << 8), x, y, z]
    return words, 4 * n_gates, len(prog)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="20,22,24")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch

    zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
    zk.init(0)
    lib, H = zk.lib(), zk.host
    chk = zk.capi.check
    st = torch.cuda.current_stream().cuda_stream
    ops, ms_ = C.c_double(), C.c_double()
    chk(lib.b200zk_microbench(7, 4000, C.byref(ops), C.byref(ms_)))
    imad_peak = ops.value
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6535.7   # fallback: the pool's measured copy bandwidth (B200_PROFILING.md)

    def timed(fn, reps=args.reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def report(op, log_n, n, ms, bytes_per, muls_per, **extra):
        gbs = bytes_per * n / (ms * 1e-3) / 1e9
        lmac = muls_per * bench.LMAC_PER_FR_MUL * n / (ms * 1e-3)
        t_hbm, t_int = bytes_per * n / (hbm_peak * 1e9), muls_per * bench.LMAC_PER_FR_MUL * n / imad_peak
        print(json.dumps({"op": op, "log_n": log_n, "ms": ms, "elements_per_s": n / ms * 1e3, "algorithmic_bytes_per_element": bytes_per,
                          "fr_products_per_element": muls_per, "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak,
                          "imad_frac": lmac / imad_peak, "binding": "hbm" if t_hbm >= t_int else "imad",
                          "frac_of_binding_bound": max(t_hbm, t_int) / (ms * 1e-3), **extra}), flush=True)

    rng = random.Random(1)
    for log_n in [int(x) for x in args.sizes.split(",")]:
        n = 1 << log_n
        mk = lambda seed: torch.from_numpy(bench.synth_scalars_np(seed, 0, n).view(np.uint8).reshape(-1)).cuda()
        a, b, out = mk(11), mk(12), torch.empty(32 * n, dtype=torch.uint8, device="cuda")
        for t in (a, b):   # Montgomery form, like resident prover columns
            chk(lib.b200zk_fr_convert_dev(t.data_ptr(), t.data_ptr(), n, 1, st))
        report("fr_convert", log_n, n, timed(lambda: chk(lib.b200zk_fr_convert_dev(a.data_ptr(), out.data_ptr(), n, 1, st))), 64, 1)
        report("fr_pointwise_mul", log_n, n, timed(lambda: chk(lib.b200zk_fr_pointwise_dev(0, a.data_ptr(), b.data_ptr(), None, out.data_ptr(), n, st))), 96, 1)
        polys = [mk(20 + i) for i in range(8)]
        ptrs = (C.c_void_p * 8)(*[p.data_ptr() for p in polys])
        cb = b"".join(fr(rng.randrange(R)) for _ in range(8))
        report("fr_lincomb_8", log_n, n, timed(lambda: chk(lib.b200zk_fr_lincomb_dev(C.addressof(ptrs), zk.capi.addr(cb), 8, out.data_ptr(), n, st))), 32 * 9, 8)
        del polys
        report("fr_batch_invert", log_n, n, timed(lambda: chk(lib.b200zk_fr_batch_invert_dev(a.data_ptr(), out.data_ptr(), n, st))), 160, 3 + 60.0 / 32,
               note="3 products per element + one branch-free inversion (~60 product-equivalents) per 32 elements; bytes = in twice, prefix write + read, out")
        report("fr_running_product", log_n, n, timed(lambda: chk(lib.b200zk_fr_running_product_dev(a.data_ptr(), out.data_ptr(), n, None, 0, st))), 96, 3,
               note="two sweeps: read, read + write; 3 products per element plus the block scans")
        zb = fr(rng.randrange(R))
        ev = torch.empty(32, dtype=torch.uint8, device="cuda")
        report("fr_kate_div", log_n, n, timed(lambda: chk(lib.b200zk_fr_kate_div_dev(a.data_ptr(), n, zk.capi.addr(zb), out.data_ptr(), ev.data_ptr(), st))), 96, 3)
        report("fr_poly_eval", log_n, n, timed(lambda: chk(lib.b200zk_fr_kate_div_dev(a.data_ptr(), n, zk.capi.addr(zb), None, ev.data_ptr(), st))), 64, 3)
        # gate program: extended domain = this size, k = log_n - 2, 20 columns, 24 gates
        n_cols, rots = 20, [0, 1, -1, 2]
        cols = [a, b] + [mk(40 + i) for i in range(n_cols - 2)]
        words, muls, n_instr = atms_like_program(n_cols, len(rots), 24, rng)
        t_inv = [rng.randrange(1, R) for _ in range(4)]
        gp = H.GateProgram(words, [rng.randrange(R), 0], rots, n_cols, log_n - 2, log_n, t_inv)
        cp = (C.c_void_p * n_cols)(*[c.data_ptr() for c in cols])
        ms = timed(lambda: chk(lib.b200zk_gate_program_run_dev(gp.handle, C.addressof(cp), out.data_ptr(), 0, st)), reps=3)
        report("gate_program", log_n, n, ms, 32 * (n_cols + 1), muls + 1, instructions=n_instr, columns=n_cols,
               note="bytes = every column once + the output (rotated re-reads hit L2); products = program multiplications + 1/(X^n-1)")
        gp.release()
        del cols
        torch.cuda.empty_cache()

    # ---- G1 side: SRS generation, fixed-base multiplication, decompression
    for k in (16, 20):
        n = 1 << k
        g = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
        gl = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
        sb = fr(0x5EC2E7 ** 5)
        wb = fr(pow(H.ROOT_OF_UNITY, 1 << (32 - k), R))
        ms = timed(lambda: chk(lib.b200zk_srs_generate_dev(zk.capi.addr(sb), k, zk.capi.addr(wb), g.data_ptr(), gl.data_ptr(), st)), reps=2)
        madds = 2 * n * 32
        print(json.dumps({"op": "srs_generate", "k": k, "ms": ms, "points_per_s": 2 * n / ms * 1e3, "mixed_additions": madds,
                          "imad_frac": madds * 3020 / (ms * 1e-3) / imad_peak,
                          "note": "g and g_lagrange: 2 * 2^k fixed-base multiplications (32 byte windows) + one inversion per point"}), flush=True)
    n = 26624 + 1024   # the right and left sums of a 1024-proof batch
    pts = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
    chk(lib.b200zk_g1_synth_bases_dev(bench.BASE_SEED, 0, n, pts.data_ptr(), st))
    torch.cuda.synchronize()
    wire = H.g1_export(type("B", (), {"ptr": pts.data_ptr()})(), n)
    comp = b"".join(H.g1_compress(wire[96 * i:96 * i + 96]) for i in range(n))
    d_comp = torch.frombuffer(bytearray(comp), dtype=torch.uint8).cuda()
    d_out = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.int32, device="cuda")
    ms = timed(lambda: chk(lib.b200zk_g1_decompress_dev(d_comp.data_ptr(), n, d_out.data_ptr(), None, d_st.data_ptr(), st)))
    assert int(d_st.abs().sum().item()) == 0 and bytes(d_out.cpu().numpy()) == bytes(pts.cpu().numpy())
    print(json.dumps({"op": "g1_decompress", "points": n, "ms": ms, "points_per_s": n / ms * 1e3,
                      "imad_frac": n * 575 * 300 / (ms * 1e-3) / imad_peak,
                      "note": "one 381-bit square root (~575 Fp products) per point; the batch is too small to fill 148 SMs x 4 CTAs"}), flush=True)
    sc = torch.from_numpy(bench.synth_scalars_np(9, 0, n).view(np.uint8).reshape(-1)).cuda()
    out = C.create_string_buffer(96)
    import time
    h_pts, h_sc = bytes(d_out.cpu().numpy()), bytes(sc.cpu().numpy())   # Montgomery points back on the host
    chk(lib.b200zk_msm_g1_adhoc(zk.capi.addr(h_pts), zk.FMT_MONT, zk.capi.addr(h_sc), 0, n, zk.capi.addr(out)))
    t0 = time.perf_counter()
    for _ in range(args.reps):
        chk(lib.b200zk_msm_g1_adhoc(zk.capi.addr(h_pts), zk.FMT_MONT, zk.capi.addr(h_sc), 0, n, zk.capi.addr(out)))
    e2e = (time.perf_counter() - t0) / args.reps * 1e3
    print(json.dumps({"op": "batch_verify_msm_adhoc", "points": n, "e2e_ms": e2e, "points_per_s": n / e2e * 1e3,
                      "note": "host points + scalars in, one G1 point out (register, MSM, release); pairing excluded"}), flush=True)


if __name__ == "__main__":
    main()
