"""Where the end-to-end time of the dendrogram path goes: phases of Job(...) + for_each_tree on a config, with the
CUDA-event times of the kernel groups (pair sums, GOM scoring, fused tail, device UPGMA).
usage: e2e_probe.py [cfg3] [replicates] [ring_matrices]"""
import ctypes as C, sys, time, numpy as np
sys.path.insert(0, ".")
import gravity2_b200
from gravity2_b200 import synth, gom_membership
from gravity2_b200.engine import Job
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ring = int(sys.argv[3]) if len(sys.argv) > 3 else 0
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
    got, first = [], []
    if rep == 1:
        eng.lib.gv2_timing_enable(eng.ctx, 1)
    job.for_each_tree(w, names, lambda b, s: (got.append(len(s)), first.append(time.perf_counter()) if not first else None),
                      ring_bytes=(ring * n * n * 8 if ring else None))
    t2 = time.perf_counter()
    kinds = {}
    if rep == 1:
        for name, kind in (("pairs", 0), ("tail", 1), ("gom", 2), ("linkage", 3)):
            ms, cnt = C.c_double(), C.c_int64()
            eng.lib.gv2_timing_read_kind(eng.ctx, kind, C.byref(ms), C.byref(cnt))
            kinds[name] = (round(ms.value, 1), cnt.value)
        eng.lib.gv2_timing_enable(eng.ctx, 0)
    job.close()
    t3 = time.perf_counter()
    print(f"pass {rep}: job create {t1 - t0:.3f} s, {len(got)} trees {t2 - t1:.3f} s (first tree after {first[0] - t1:.3f} s), close {t3 - t2:.3f} s, "
          f"total {t3 - t0:.3f} s; kernel groups (ms, launches): {kinds}")
