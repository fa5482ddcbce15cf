import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
from gravity2_b200 import _lib
lib = _lib.load()
ctx = C.c_void_p(); lib.gv2_create_on_stream(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(ctx))
def run(n, p, nb, weighted):
    n_pad = (n+127)//128*128; p_pad = (p+31)//32*32
    nt = int(lib.gv2_num_tiles(n))
    tab = torch.rand((n_pad, p_pad), dtype=torch.float32, device="cuda")
    wf = torch.ones((nb, p_pad), dtype=torch.float32, device="cuda")
    ws = torch.empty((nb, nt, 16384), dtype=torch.float64, device="cuda")
    f = lambda: lib.gv2_dev_gj_min_sums(ctx, C.c_void_p(tab.data_ptr()), n_pad, p_pad, 0, nt, C.c_void_p(wf.data_ptr()) if weighted else None, nb, 1, C.c_void_p(ws.data_ptr()))
    for _ in range(2): assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/3
    cells = nt*16384.0*p_pad*nb
    print(f"n={n} p={p} nb={nb} w={weighted}: {ms:.2f} ms, {cells/ms/1e9:.2f} Tcell/s processed = {cells/ms/1e9/18.6*100:.1f}% of 18.6")
run(1024, 4992, 100, True); run(1024, 4992, 100, False) if False else None
run(1024, 4992, 1, False)
run(4096, 30016, 1, False); run(4096, 30016, 2, True)
