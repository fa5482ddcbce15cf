"""A scaled-down cfg3-like PG job (same per-CTA behaviour, a few GB of device memory) for `ncu --set full` captures:
ncu saves and restores the device memory a kernel writes around each of its ~40 replays, so the full cfg3 job (100 GB
workspaces) must never be profiled this way.

usage: prof_small.py [N=4096] [P=30000] [G=80] [NB=64] [density=0.01]
Prints the per-group CUDA-event times of one step (base + NB replicates, device-only)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gravity2_b200 import _lib, gom_membership, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
g = int(sys.argv[3]) if len(sys.argv) > 3 else 80
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 64
dens = float(sys.argv[5]) if len(sys.argv) > 5 else 0.01
lib = _lib.load()
ctx = C.c_void_p()
assert lib.gv2_create_on_stream(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(ctx)) == 0
t = synth.make_tables_sparse(n, p, g, dens, seed=7)
gor = gom_membership(t.taxo, t.gom_ids)
d_sig = torch.from_numpy(t.sig).cuda()
rng = np.random.default_rng(1)
w = torch.from_numpy(np.stack([np.bincount(rng.integers(0, p, p), minlength=p) for _ in range(nb)]).astype(np.int32)).cuda()
job = C.c_void_p()
rc = lib.gv2_job_create(ctx, C.c_void_p(d_sig.data_ptr()), C.c_void_p(d_sig.data_ptr()), 1, n, p, gor.ctypes.data_as(C.c_void_p), n, len(t.gom_ids),
                        _lib.SCHEMES["PG"], 1.0, _lib.OUT_DISTANCE, C.byref(job))
assert rc == 0, lib.gv2_last_error(ctx)
ring = max(1, min(nb, int((2 << 30) // (n * n * 8))))
out = torch.empty((ring, n, n), dtype=torch.float64, device="cuda")
base = torch.empty((n, n), dtype=torch.float64, device="cuda")
ntiles = int(lib.gv2_num_tiles(n))


def step():
    assert lib.gv2_job_base(job, 0, ntiles, 0, C.c_void_p(base.data_ptr())) == 0, lib.gv2_last_error(ctx)
    rc = lib.gv2_job_replicates_stream(job, C.c_void_p(w.data_ptr()), nb, C.c_void_p(out.data_ptr()), ring, None, None, None)
    assert rc == 0, lib.gv2_last_error(ctx)


step()
torch.cuda.synchronize()
lib.gv2_timing_enable(ctx, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
line = [f"n={n} p={p} g={len(t.gom_ids)} nb={nb}: step {e0.elapsed_time(e1):.2f} ms"]
for name, kind in (("pairs", _lib.TIME_PAIRSUMS), ("tail", _lib.TIME_TAIL), ("gom", _lib.TIME_GOM), ("base", _lib.TIME_BASE)):
    ms, cnt = C.c_double(0.0), C.c_int64(0)
    lib.gv2_timing_read_kind(ctx, kind, C.byref(ms), C.byref(cnt))
    line.append(f"{name} {ms.value:.2f} ms / {cnt.value}")
print("; ".join(line), f"; device bytes {lib.gv2_job_device_bytes(job) / 1e9:.2f} GB")
lib.gv2_job_destroy(job)
