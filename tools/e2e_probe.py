"""Where the end-to-end time of the dendrogram path goes: phases of Job(...) + for_each_tree on a config.
usage: e2e_probe.py [cfg3] [replicates]"""
import ctypes as C, sys, time, numpy as np
sys.path.insert(0, ".")
import gravity2_b200
from gravity2_b200 import synth, gom_membership
from gravity2_b200.engine import Job
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
t = synth.make_config(cfg, fast=True)
n, p = t.sig.shape
idx = synth.reference_resample_indices(p, nb, seed=1)
w = np.concatenate([np.ones((1, p), np.int32), np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32)])
gor = gom_membership(t.taxo, t.gom_ids)
names = [f"{x}_{i}" for i, x in enumerate(t.taxo)]
eng = gravity2_b200.default_engine()
for rep in range(2):
    t0 = time.perf_counter()
    job = Job(eng, t.sig, t.sig, gor, len(t.gom_ids), "PG", 1.0, distance=True)
    eng.synchronize()
    t1 = time.perf_counter()
    got = []
    first = []
    job.for_each_tree(w, names, lambda b, s: (got.append(len(s)), first.append(time.perf_counter()) if not first else None))
    t2 = time.perf_counter()
    job.close()
    t3 = time.perf_counter()
    print(f"pass {rep}: job create {t1 - t0:.3f} s, {len(got)} trees {t2 - t1:.3f} s (first tree after {first[0] - t1:.3f} s), close {t3 - t2:.3f} s, total {t3 - t0:.3f} s")
