"""Single-kernel driver for ncu: a few launches of the Jaccard tile kernel (weighted + unweighted)."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from gravity2_b200 import _lib
lib = _lib.load()
ctx = C.c_void_p(); lib.gv2_create_on_stream(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(ctx))
n, p, nb = 2048, 8192, 4
n_pad, p_pad = n, p
nt = int(lib.gv2_num_tiles(n))
tab = torch.rand((n_pad, p_pad), dtype=torch.float32, device="cuda")
wf = torch.ones((nb, p_pad), dtype=torch.float32, device="cuda")
ws = torch.empty((nb, nt, 16384), dtype=torch.float64, device="cuda")
for weighted in (False, True, False, True):
    rc = lib.gv2_dev_gj_min_sums(ctx, C.c_void_p(tab.data_ptr()), n_pad, p_pad, 0, nt, C.c_void_p(wf.data_ptr()) if weighted else None, nb if weighted else 1, 1, C.c_void_p(ws.data_ptr()))
    assert rc == 0
torch.cuda.synchronize()
print("ok")
