"""Time the GOM database build and the (weighted) GOM table kernels for a synthetic config."""
import ctypes as C, sys, time, numpy as np, torch
sys.path.insert(0, ".")
from gravity2_b200 import _lib, synth, gom_membership
n, p, g, dens, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
t0 = time.time(); t = synth.make_tables(n, p, g, dens, seed=3); print(f"tables {time.time()-t0:.1f}s nnz/row={np.count_nonzero(t.sig)/n:.0f}")
lib = _lib.load()
ctx = C.c_void_p(); lib.gv2_create_on_stream(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(ctx))
gor = gom_membership(t.taxo, t.gom_ids)
d_sig = torch.from_numpy(t.sig).cuda()
job = C.c_void_p()
torch.cuda.synchronize(); t0 = time.time()
rc = lib.gv2_job_create(ctx, C.c_void_p(d_sig.data_ptr()), C.c_void_p(d_sig.data_ptr()), 1, n, p, gor.ctypes.data_as(C.c_void_p), n, g, 2, 1.0, 1, C.byref(job)); assert rc == 0, lib.gv2_last_error(ctx)
torch.cuda.synchronize(); print(f"job_create {time.time()-t0:.2f}s, device bytes {lib.gv2_job_device_bytes(job)/1e9:.2f} GB")
w = torch.poisson(torch.ones((nb, p), device="cuda")).to(torch.int32)
out = torch.empty((nb, n, g), dtype=torch.float64, device="cuda")
for label, wp, k in (("base", None, 1), (f"{nb} replicates", C.c_void_p(w.data_ptr()), nb)):
    lib.gv2_job_gom_table(job, wp, k, C.c_void_p(out.data_ptr())); torch.cuda.synchronize()
    t0 = time.time(); rc = lib.gv2_job_gom_table(job, wp, k, C.c_void_p(out.data_ptr())); torch.cuda.synchronize()
    assert rc == 0, lib.gv2_last_error(ctx)
    print(f"gom_table {label}: {(time.time()-t0)*1e3:.1f} ms -> {(time.time()-t0)*1e3/k:.2f} ms/replicate; nan={int(torch.isnan(out[:k]).sum())} mean={float(out[:k].nan_to_num().mean()):.4f}")
