"""Phases of the single-process multi-GPU dendrogram path (MultiEngine / MultiJob): job creation, trees, close.
usage: multi_probe.py [cfg3] [replicates] [n_gpus]"""
import sys, time, numpy as np
sys.path.insert(0, ".")
import gravity2_b200 as G
from gravity2_b200 import synth, gom_membership
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ng = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t = synth.make_config(cfg, fast=True)
n, p = t.sig.shape
idx = synth.reference_resample_indices(p, nb, seed=1)
w = np.concatenate([np.ones((1, p), np.int32), np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32)])
gor = gom_membership(t.taxo, t.gom_ids)
names = [f"{x}_{i}" for i, x in enumerate(t.taxo)]
t0 = time.perf_counter()
meng = G.MultiEngine(list(range(ng)))
print(f"MultiEngine({ng}): {time.perf_counter() - t0:.3f} s, nccl={meng.uses_nccl}")
for rep in range(2):
    t0 = time.perf_counter()
    job = G.MultiJob(meng, t.sig, t.sig, gor, len(t.gom_ids), "PG", 1.0, distance=True)
    t1 = time.perf_counter()
    got, first = [], []
    job.for_each_tree(w, names, lambda b, s: (got.append(len(s)), first.append(time.perf_counter()) if not first else None))
    t2 = time.perf_counter()
    job.close()
    t3 = time.perf_counter()
    print(f"pass {rep}: job create {t1 - t0:.3f} s, {len(got)} trees {t2 - t1:.3f} s (first after {first[0] - t1:.3f} s), close {t3 - t2:.3f} s, "
          f"total {t3 - t0:.3f} s")
meng.close()
