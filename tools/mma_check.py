"""tcgen05 bootstrap kernel vs the CUDA-core weighted kernel vs the fp64 oracle (scheme P)."""
import os, sys, time, numpy as np
sys.path.insert(0, ".")
from gravity2_b200 import synth, Engine
from oracle import gravity_oracle as O
n, p, nb, dens = [int(x) if i < 3 else float(x) for i, x in enumerate((sys.argv[1:] + ["200", "700", "40", "0.3"])[:4])]
t = synth.make_tables(n, p, 5, dens, seed=5)
rng = np.random.default_rng(1)
w = np.stack([np.bincount(rng.integers(0, p, p), minlength=p) for _ in range(nb)]).astype(np.int32)
eng = Engine(0)
t0 = time.time(); S = eng.bootstrap(t.sig, None, None, 0, w, "P", 1.0, distance=False); print("device path", time.time() - t0, "s, launches", eng.launches)
worst = 0
for r in sorted(set([0, 1, nb // 2, nb - 1])):
    ref = O.similarity_mat_fast(t.sig, None, None, 0, "P", 1.0, w=w[r])
    rel = np.abs(S[r] - ref) / np.maximum(ref, 1e-300)
    rel[ref == 0] = np.abs(S[r])[ref == 0]
    worst = max(worst, rel.max())
    print("replicate", r, "max rel err", rel.max(), "sym", np.array_equal(S[r], S[r].T), "diag1", np.all(np.diag(S[r])[t.sig.any(1)] == 1.0))
print("MMA" if os.environ.get("GV2_DISABLE_MMA") != "1" else "CUDA-core", "worst", worst)
assert worst < 1e-6
