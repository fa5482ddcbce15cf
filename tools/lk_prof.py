"""Device linkage of one 1 - Jaccard matrix, kernel time by variant (GV2_LK_LAZY / GV2_LK_LIST are read at every launch);
with one variant on the command line it is the ncu target: python tools/lk_prof.py [n] [method] [variant]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gravity2_b200 as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
method = sys.argv[2] if len(sys.argv) > 2 else "average"
only = sys.argv[3] if len(sys.argv) > 3 else None
rng = np.random.default_rng(0)
a = rng.random((n, 64)) < 0.15
a[:, 0] = True
inter = (a.astype(np.float32) @ a.astype(np.float32).T).astype(np.float64)
cnt = a.sum(1).astype(np.float64)
D = 1.0 - inter / (cnt[:, None] + cnt[None, :] - inter)
np.fill_diagonal(D, 0.0)
eng = G.default_engine()
eng.linkage(D[:64, :64].copy(), method)
eng.lib.gv2_timing_enable(eng.ctx, 1)
variants = {"eager": {"GV2_LK_LAZY": "0"}, "lazy": {"GV2_LK_LIST": "0"}, "list": {}}
ref = None
for name, env in variants.items():
    if only and name != only:
        continue
    for k in ("GV2_LK_LAZY", "GV2_LK_LIST"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for rep in range(2):
        ms0, c0 = C.c_double(), C.c_int64()
        eng.lib.gv2_timing_read_kind(eng.ctx, 3, C.byref(ms0), C.byref(c0))
        t0 = time.time()
        Z = eng.linkage(D, method)
        wall = time.time() - t0
        ms1, c1 = C.c_double(), C.c_int64()
        eng.lib.gv2_timing_read_kind(eng.ctx, 3, C.byref(ms1), C.byref(c1))
        print(f"n={n} {method} {name}: kernel {ms1.value:.1f} ms, call {wall:.3f} s", flush=True)   # (a read resets the sum)
    ref = Z if ref is None else ref
    print("  identical to the first variant:", np.array_equal(Z, ref), flush=True)
