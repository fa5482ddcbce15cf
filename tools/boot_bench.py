"""Time gv2_job_replicates (scheme P) for a given shape with and without the tensor-core path."""
import ctypes as C, os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
from gravity2_b200 import _lib
n, p, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lib = _lib.load()
ctx = C.c_void_p(); lib.gv2_create_on_stream(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(ctx))
g = torch.Generator(device="cuda"); g.manual_seed(1)
sig = (torch.rand((n, p), generator=g, device="cuda", dtype=torch.float64) * 500 + 0.1) * (torch.rand((n, p), generator=g, device="cuda") < 0.02)
w = torch.poisson(torch.ones((nb, p), device="cuda")).to(torch.int32)
job = C.c_void_p()
rc = lib.gv2_job_create(ctx, C.c_void_p(sig.data_ptr()), None, 1, n, p, None, 0, 0, 0, 1.0, 1, C.byref(job)); assert rc == 0, lib.gv2_last_error(ctx)
ring = max(1, min(nb, int((8 << 30) // (n * n * 8))))
out = torch.empty((ring, n, n), dtype=torch.float64, device="cuda")
def step():
    rc = lib.gv2_job_replicates_stream(job, C.c_void_p(w.data_ptr()), nb, C.c_void_p(out.data_ptr()), ring, None, None, None); assert rc == 0, lib.gv2_last_error(ctx)
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
pairs = n * (n + 1) // 2
print(f"n={n} p={p} nb={nb} ring={ring} mma={'off' if os.environ.get('GV2_DISABLE_MMA')=='1' else 'on'}: {ms:.1f} ms, {pairs*nb/ms/1e6:.2f} G pair-sims/s, "
      f"{pairs*nb*p/ms/1e9:.1f} T cell-rep/s, bytes={lib.gv2_job_device_bytes(job)/1e9:.1f} GB")
