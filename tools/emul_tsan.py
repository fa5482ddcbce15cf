"""ThreadSanitizer over the emulated kernels: builds every tests/emul driver with -fsanitize=thread and runs one
representative case each; a CUDA block is one OS thread per CUDA thread with barriers for __syncthreads / warp primitives
(tests/emul/cuda_block_emul.h), so a data race in a kernel's shared-memory or global-memory protocol inside a block shows up
as a TSan report.  Prints, per case, the exit code and the distinct (access, file:line) pairs reported.
usage: python tools/emul_tsan.py > profiles/rNN_emul_tsan.log"""
import collections
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gravity2_b200 import synth  # noqa: E402
from oracle import gravity_oracle as O  # noqa: E402

tmp = tempfile.mkdtemp()
exe = {}
for name in ("lk_emul", "popc_emul", "small_emul", "gj_emul", "bs_emul", "gom_emul", "mma_emul"):
    exe[name] = os.path.join(tmp, name)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-ffp-contract=off", "-fsanitize=thread", "-o", exe[name],
                    os.path.join(ROOT, "tests", "emul", name + ".cpp")], check=True, stderr=subprocess.DEVNULL)


def run(label, name, args, blob):
    # cp.async copies land at issue here ("early"): a prefetch into a stage that is still being read is then a reported race
    r = subprocess.run([exe[name]] + [str(a) for a in args], input=blob, capture_output=True, timeout=3600,
                       env=dict(os.environ, GV2_EMUL_CP_ASYNC="early"))
    txt = r.stderr.decode(errors="replace")
    pairs = collections.Counter()
    for blk in txt.split("=================="):
        ls = re.findall(r"((?:Previous )?(?:[Aa]tomic )?(?:[Rr]ead|[Ww]rite)) of size (\d+).*?\n\s+#0 .*?(\w+\.(?:cuh|h|cpp)):(\d+)", blk)
        if ls:
            pairs[" / ".join(f"{a} {f}:{l}" for a, _, f, l in ls)] += 1
    print(f"{label}: exit {r.returncode}, reports: {dict(pairs) if pairs else 'none'}", flush=True)


rng = np.random.default_rng(3)
n = 130
D = rng.integers(1, 6, (n, n)) / 5.0
D = np.triu(D, 1)
D = D + D.T
for method, code, variant in (("average, eager columns", 0, 0), ("average, lazy columns", 0, 1), ("average, live list", 0, 2), ("ward, live list", 3, 2),
                              ("single", 4, 0), ("centroid", 5, 0), ("median", 6, 0)):
    run(f"linkage kernels ({method}), n = {n}", "lk_emul", [code, n, variant], D.tobytes())
tab = np.where(rng.random((130, 1025)) < 0.1, 1.5, 0.0)
for packed in (0, 1):
    run(f"presence/absence kernels, 130 x 1025, packed = {packed}", "popc_emul", [130, 1025, packed], tab.tobytes())
S = np.round(rng.random((9, 700)), 2)
run("rowmax_block_kernel", "small_emul", ["rowmax", 9, 700, 2, 9, 13, 650], S.tobytes())
run("philox_indices_kernel", "small_emul", ["philox", 20241019, 0, 3, 1031], b"")
t = synth.make_tables(140, 300, 6, 0.04, seed=11)
sig = t.sig.copy()
sig[17, 5] = np.nan
gom = rng.random((140, 70)) * (rng.random((140, 70)) < 0.7)
run("prepare + K1 (split K) + K3, scheme PG, 140 x 300", "gj_emul", ["similarity", 140, 300, 70, 2, 1.0, 1, 0.0, 1, 2], sig.tobytes() + gom.tobytes())
run("fused tail, fixed point, scheme PG", "gj_emul", ["fused", 140, 300, 70, 1, 1.0, 1], sig.tobytes() + gom.tobytes())
run("fused tail, fp32, scheme G, powered", "gj_emul", ["fused", 140, 0, 70, 0, 2.0, 0], gom.tobytes())
t2 = synth.make_tables(130, 200, 5, 0.08, seed=1)
for nb in (40, 3):
    w = np.stack([np.bincount(rng.integers(0, 200, 200), minlength=200) for _ in range(nb)]).astype(np.int32)
    run(f"sparse pair sums (lists + {'stream' if nb > 12 else 'pair-lane'} kernel), 130 x 200, {nb} replicates", "bs_emul", [130, 200, nb],
        t2.sig.tobytes() + w.tobytes())
t3 = synth.make_tables(20, 300, 3, 0.3, seed=8)
w = np.stack([np.bincount(rng.integers(0, 300, 300), minlength=300) for _ in range(40)]).astype(np.int32)
db = O.gomdb(t3.taxo, t3.loc, t3.gom_ids)
rows = [np.asarray(db[k]).reshape(-1, 300) for k in t3.gom_ids]
offs = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
run("GOM scoring (build + 40 replicates, light and heavy pairs), 20 x 300", "gom_emul", [20, 300, 3, 40, 0, 1],
    t3.loc.tobytes() + offs.tobytes() + np.concatenate(rows).tobytes() + w.tobytes())
t4 = synth.make_tables(24, 520, 4, 0.3, seed=5)
w = np.stack([np.bincount(rng.integers(0, 520, 520), minlength=520) for _ in range(20)]).astype(np.int32)
run("tensor-core pair sums (producers / TMA / MMA warps over the tcgen05 model), 24 x 520, 20 replicates", "mma_emul", [24, 520, 20],
    t4.sig.tobytes() + w.tobytes())
