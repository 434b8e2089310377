import ctypes as C, importlib, sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
zk = importlib.import_module("plutus-halo2-verifier-gen_b200")
zk.init(0)
lib, chk = zk.lib(), zk.capi.check
n = 1 << 24
st = torch.cuda.current_stream().cuda_stream
d_b = torch.empty(96 * n, dtype=torch.uint8, device="cuda")
chk(lib.b200zk_g1_synth_bases_dev(bench.BASE_SEED, 0, n, d_b.data_ptr(), st)); torch.cuda.synchronize()
h = C.c_uint64(0)
chk(lib.b200zk_bases_register_dev(d_b.data_ptr(), n, zk.FMT_MONT, 96, C.byref(h)))
del d_b
sc = bench.synth_scalars_np(1, 0, n)          # ordinary (pageable) numpy memory, like a Rust Vec
out = C.create_string_buffer(96)
for _ in range(2):
    chk(lib.b200zk_msm_g1(h.value, 0, sc.ctypes.data, n, 0, zk.capi.addr(out)))
t0 = time.perf_counter()
for _ in range(3):
    chk(lib.b200zk_msm_g1(h.value, 0, sc.ctypes.data, n, 0, zk.capi.addr(out)))
print("pageable e2e ms", (time.perf_counter() - t0) / 3 * 1e3, zk.host.g1_compress(out.raw).hex()[:16])
