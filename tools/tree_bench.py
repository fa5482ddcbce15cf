"""Time Job.for_each_tree (replicate dendrograms, matrices stay in HBM) on a synthetic config.
usage: tree_bench.py [cfg] [replicates]"""
import ctypes as C, sys, time, numpy as np
sys.path.insert(0, ".")
import gravity2_b200
from gravity2_b200 import synth, _lib, gom_membership
from gravity2_b200.engine import Job
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
t = synth.make_config(cfg)
n, p = t.sig.shape
idx = synth.reference_resample_indices(p, nb, seed=1)
w = np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32)
eng = gravity2_b200.default_engine()
job = Job(eng, t.sig, t.sig, gom_membership(t.taxo, t.gom_ids), len(t.gom_ids), "PG", 1.0, distance=True)
names = [f"{x}_{i}" for i, x in enumerate(t.taxo)]
got = []
eng.lib.gv2_timing_enable(eng.ctx, 1)
t0 = time.perf_counter()
job.for_each_tree(w, names, lambda b, s: got.append(len(s)))
dt = time.perf_counter() - t0
out = {}
for name, kind in (("pairs", 0), ("tail", 1), ("gom", 2), ("linkage", 3)):
    ms, cnt = C.c_double(), C.c_int64()
    eng.lib.gv2_timing_read_kind(eng.ctx, kind, C.byref(ms), C.byref(cnt))
    out[name] = (round(ms.value, 1), cnt.value)
print(f"{cfg}: {nb} replicate trees of {n} leaves in {dt:.2f} s ({dt / nb * 1e3:.1f} ms each); newick {got[0]} chars; kernel groups ms, launches: {out}")
job.close()
