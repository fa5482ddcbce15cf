"""Device linkage of one matrix per method: time, and equality with scipy where scipy finishes in seconds.
python tools/linkage_bench.py [n_check] [n_time]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gravity2_b200 as G

n_check = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
eng = G.default_engine()
rng = np.random.default_rng(0)


def matrix(n):
    a = rng.random((n, 64)) < 0.15                                   # 1 - binary Jaccard: the shape of the real distances
    a[:, 0] = True
    inter = (a.astype(np.float32) @ a.astype(np.float32).T).astype(np.float64)
    cnt = a.sum(1).astype(np.float64)
    D = 1.0 - inter / (cnt[:, None] + cnt[None, :] - inter)
    np.fill_diagonal(D, 0.0)
    return D


from scipy.cluster.hierarchy import linkage
D = matrix(n_check)
for m in ("average", "centroid", "median"):
    t0 = time.time(); Z = eng.linkage(D, m); t1 = time.time()
    with np.errstate(all="ignore"):
        ref = linkage(D[np.triu_indices(n_check, 1)], m)
    t2 = time.time()
    print(f"n={n_check} {m}: device {t1 - t0:.3f} s (incl. upload), scipy {t2 - t1:.3f} s, identical={np.array_equal(Z, ref, equal_nan=True)}", flush=True)
D = matrix(n_time)
for m in ("average", "centroid", "median"):
    eng.linkage(D[:64, :64].copy(), m)
    t0 = time.time(); Z = eng.linkage(D, m); t1 = time.time()
    print(f"n={n_time} {m}: device {t1 - t0:.3f} s (incl. 0.8 GB upload)", flush=True)
