"""The centroid / median linkage kernel, run on the CPU: tests/emul/cuda_block_emul.h supplies one CUDA thread block as host
threads (a barrier for __syncthreads, per-warp barriers for the REDUX reductions), tests/emul/lk_generic_emul.cpp includes the
SAME source the GPU build compiles (gravity2_b200/csrc/linkage_generic.cuh).  What is checked here is the kernel's logic
-- barrier placement, the hand-overs through shared memory, the order of the heap operations that decides how ties fall --
against scipy; the GPU tests (test_gpu_parity.py::test_device_linkage_equals_scipy) check the compiled kernel itself."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    exe = str(tmp_path_factory.mktemp("emul") / "lk_generic_emul")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-ffp-contract=off", "-o", exe,
                    os.path.join(ROOT, "tests", "emul", "lk_generic_emul.cpp")], check=True)
    return exe


def _matrix(kind, n, rng):
    if kind == "euclid":
        x = rng.random((n, 4))
        D = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
    elif kind == "random":                                       # not a metric: square roots of negative numbers on the way
        D = rng.random((n, n))
    elif kind == "ties":
        D = rng.integers(1, 6, (n, n)) / 5.0
    elif kind == "blocks":                                       # taxon-like: many exact 1.0
        grp = rng.integers(0, 6, n)
        D = np.where(grp[:, None] == grp[None, :], np.round(rng.random((n, n)) * 0.6, 2), 1.0)
    else:                                                        # 1 - binary Jaccard, the sort_pphmm distances
        a = rng.random((n, 30)) < 0.2
        a[:, 0] = True
        D = 1 - (a[:, None, :] & a[None, :, :]).sum(-1) / (a[:, None, :] | a[None, :, :]).sum(-1)
    D = np.triu(D, 1)
    return D + D.T


@pytest.mark.parametrize("method,code", [("centroid", 5), ("median", 6)])
@pytest.mark.parametrize("kind,n", [("euclid", 2), ("ties", 3), ("euclid", 97), ("random", 140), ("ties", 129), ("blocks", 260),
                                    ("jaccard", 75)])
def test_emulated_generic_linkage_kernel_equals_scipy(emul, kind, n, method, code):
    from scipy.cluster.hierarchy import linkage
    D = _matrix(kind, n, np.random.default_rng(n))
    with np.errstate(all="ignore"):
        ref = linkage(D[np.triu_indices(n, 1)], method)
    out = subprocess.run([emul, str(code), str(n)], input=np.ascontiguousarray(D).tobytes(), capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr
    Z = np.frombuffer(out.stdout, dtype=np.float64).reshape(n - 1, 4)
    assert np.array_equal(Z, ref, equal_nan=True)
