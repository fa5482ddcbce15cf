"""One device linkage of a 1 - Jaccard matrix (for ncu): python tools/lk_prof.py [n] [method]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gravity2_b200 as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
method = sys.argv[2] if len(sys.argv) > 2 else "average"
rng = np.random.default_rng(0)
a = rng.random((n, 64)) < 0.15
a[:, 0] = True
inter = (a.astype(np.float32) @ a.astype(np.float32).T).astype(np.float64)
cnt = a.sum(1).astype(np.float64)
D = 1.0 - inter / (cnt[:, None] + cnt[None, :] - inter)
np.fill_diagonal(D, 0.0)
eng = G.default_engine()
eng.linkage(D[:64, :64].copy(), method)
t0 = time.time()
Z = eng.linkage(D, method)
print(f"n={n} {method}: {time.time() - t0:.3f} s, last height {Z[-1, 2]:.6f}")
