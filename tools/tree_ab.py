"""A/B of the UPGMA kernels inside the bootstrap job's dendrogram mode: the same job clusters the same replicates with
GV2_LK_LIST=0 (lazy chain kernel) and =1 (lazy + live list); times, and the Newick strings must be identical.
usage: tree_ab.py [cfg] [replicates] [rounds]"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gravity2_b200
from gravity2_b200 import synth, gom_membership
from gravity2_b200.engine import Job

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 252
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t = synth.make_config(cfg)
n, p = t.sig.shape
idx = synth.reference_resample_indices(p, nb, seed=1)
w = np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32)
eng = gravity2_b200.default_engine()
job = Job(eng, t.sig, t.sig, gom_membership(t.taxo, t.gom_ids), len(t.gom_ids), "PG", 1.0, distance=True)
names = [f"{x}_{i}" for i, x in enumerate(t.taxo)]
digests = {}
for r in range(rounds):
    for mode in ("0", "1"):
        os.environ["GV2_LK_LIST"] = mode
        h = hashlib.sha256()
        t0 = time.perf_counter()
        job.for_each_tree(w, names, lambda b, s: h.update(s.encode() if isinstance(s, str) else s))
        dt = time.perf_counter() - t0
        digests.setdefault(mode, h.hexdigest())
        assert digests[mode] == h.hexdigest()
        print(f"{cfg} GV2_LK_LIST={mode}: {nb} trees of {n} leaves in {dt:.3f} s ({dt / nb * 1e3:.2f} ms each)", flush=True)
print("identical Newick output:", digests["0"] == digests["1"])
job.close()
sys.exit(0 if digests["0"] == digests["1"] else 1)
