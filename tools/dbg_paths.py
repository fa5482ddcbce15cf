import os, sys, numpy as np
sys.path.insert(0, ".")
import gravity2_b200
from gravity2_b200 import synth
from oracle import gravity_oracle as O
eng = gravity2_b200.default_engine()
n, p, dens, nb = 300, 900, 0.03, 600
t = synth.make_tables(n, p, 5, dens, seed=n + nb)
t.sig[7] = 0.0; t.sig[n - 1] = t.sig[2]
rng = np.random.default_rng(nb)
w = np.stack([np.bincount(rng.integers(0, p, p), minlength=p) for _ in range(nb)]).astype(np.int32)
for path in ("sparse", "mma", "fp32"):
    os.environ["GV2_BOOT_PATH"] = path
    out = eng.bootstrap(t.sig, None, None, 0, w, "P", 1.0, distance=False)
    for b in (0, nb // 2, nb - 1):
        ref = O.similarity_mat_fast(t.sig, None, None, 0, "P", 1.0, w=w[b])
        err = np.abs(out[b] - ref); rel = err / np.maximum(np.abs(ref), 1e-300)
        rel[err <= 1e-12] = 0
        i, j = np.unravel_index(np.argmax(rel), rel.shape)
        print(path, b, "max rel", rel.max(), "at", (i, j), out[b][i, j], ref[i, j], "nbad(1e-6)", int((rel > 1e-6).sum()))
