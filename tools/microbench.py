"""Integer-pipe micro-benchmarks on the bound B200 (the measured IMAD peak that the MSM / NTT
roofline fractions are quoted against).  Usage: python tools/microbench.py [iters]"""
import ctypes as C
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("plutus-halo2-verifier-gen_b200")


def run(iters=2000):
    zk.init(-1)
    names = {0: "imad_wide_lmac_per_s", 2: "fp_mul_per_s", 3: "xyzz_madd_per_s", 4: "fr_mul_per_s",
             5: "imad_wide_carry_chain_lmac_per_s", 6: "dfma_per_s", 7: "imad_wide_carry_out_lmac_per_s", 8: "imad_wide_plus_iadd_lmac_per_s",
             9: "imad_hi_only_per_s", 10: "imad_lo_only_per_s", 11: "imad_unfused_pair_imm_lmac_per_s",
             12: "fr_mul_16warps_per_s", 13: "fr_butterfly_16warps_per_s", 14: "fr_butterfly_64warps_per_s"}
    out = {"device": zk.device_info()}
    for kind, name in names.items():
        ops, ms = C.c_double(), C.c_double()
        it = iters * (1 if kind in (2, 3, 4, 12) else (1 if kind in (13, 14) else 8)) // (4 if kind in (13, 14) else 1)
        zk.capi.check(zk.lib().b200zk_microbench(kind, it, C.byref(ops), C.byref(ms)))
        out[name] = ops.value
        out[name.replace("_per_s", "_ms")] = ms.value
    return out


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 2000)))
