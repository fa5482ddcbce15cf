"""Time gv2_job_replicates (scheme P) for a given shape; GV2_BOOT_PATH=sparse|mma|fp32 picks the replicate path.
usage: boot_bench.py N P NB [density] [synth]   (synth: the taxon-structured generator of gravity2_b200.synth)"""
import ctypes as C, os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
from gravity2_b200 import _lib, synth
n, p, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dens = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
lib = _lib.load()
ctx = C.c_void_p(); lib.gv2_create_on_stream(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(ctx))
g = torch.Generator(device="cuda"); g.manual_seed(1)
if len(sys.argv) > 5:
    sig = torch.from_numpy(synth.make_tables(n, p, max(1, n // 50), dens, seed=5).sig).cuda()
else:
    sig = (torch.rand((n, p), generator=g, device="cuda", dtype=torch.float64) * 500 + 0.1) * (torch.rand((n, p), generator=g, device="cuda") < dens)
w = torch.poisson(torch.ones((nb, p), device="cuda")).to(torch.int32)
job = C.c_void_p()
rc = lib.gv2_job_create(ctx, C.c_void_p(sig.data_ptr()), None, 1, n, p, None, 0, 0, 0, 1.0, 1, C.byref(job)); assert rc == 0, lib.gv2_last_error(ctx)
ring = max(1, min(nb, int((8 << 30) // (n * n * 8))))
out = torch.empty((ring, n, n), dtype=torch.float64, device="cuda")
def step():
    rc = lib.gv2_job_replicates_stream(job, C.c_void_p(w.data_ptr()), nb, C.c_void_p(out.data_ptr()), ring, None, None, None); assert rc == 0, lib.gv2_last_error(ctx)
step(); torch.cuda.synchronize()
lib.gv2_timing_enable(ctx, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
kms, kl = C.c_double(), C.c_int64(); lib.gv2_timing_read(ctx, C.byref(kms), C.byref(kl))
pairs = n * (n + 1) // 2
print(f"n={n} p={p} nb={nb} dens={dens} path={os.environ.get('GV2_BOOT_PATH','auto')}: {ms:.1f} ms (replicate kernel {kms.value:.1f} ms / {kl.value} launches), "
      f"{pairs*nb/ms/1e6:.2f} G pair-sims/s, bytes={lib.gv2_job_device_bytes(job)/1e9:.1f} GB")
