"""The linkage kernels, run on the CPU: tests/emul/cuda_block_emul.h supplies one CUDA thread block as host threads (a barrier
for __syncthreads, per-warp barriers for the REDUX reductions and shuffles), tests/emul/lk_emul.cpp includes the SAME source
the GPU build compiles (gravity2_b200/csrc/linkage_kernels.cuh).  What is checked here is the kernels' logic -- barrier
placement, the hand-overs through shared memory, the lazy column bookkeeping, the order of the comparisons and heap
operations that decides how ties fall -- against scipy; the GPU tests (test_gpu_parity.py::test_device_linkage_equals_scipy)
check the compiled kernels themselves.  (Built with -fsanitize=thread the emulation doubles as a race detector for the
kernels' shared-memory protocol: see DESIGN.md section 4.)"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODES = {"average": 0, "complete": 1, "weighted": 2, "ward": 3, "single": 4, "centroid": 5, "median": 6}


@pytest.fixture(scope="module")
def emul(emul_build):
    return emul_build("lk_emul")


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
    elif kind == "line":                                         # the chain grows to n elements
        pos = np.cumsum(1.0 + (n - np.arange(n)) * 1e-3)
        D = np.abs(pos[:, None] - pos[None, :])
    else:                                                        # 1 - binary Jaccard, the sort_pphmm distances
        a = rng.random((n, 30)) < 0.2
        a[:, 0] = True
        D = 1 - (a[:, None, :] & a[None, :, :]).sum(-1) / (a[:, None, :] | a[None, :, :]).sum(-1)
    D = np.triu(D, 1)
    return D + D.T


def _run(emul, method, D, lazy):
    n = len(D)
    out = subprocess.run([emul, str(CODES[method]), str(n), str(int(lazy))], input=np.ascontiguousarray(D).tobytes(),
                         capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr
    return np.frombuffer(out.stdout, dtype=np.float64).reshape(n - 1, 4).copy()


@pytest.mark.parametrize("method", ["centroid", "median"])
@pytest.mark.parametrize("kind,n", [("euclid", 2), ("ties", 3), ("euclid", 97), ("random", 140), ("ties", 129), ("blocks", 260),
                                    ("jaccard", 75)])
def test_emulated_generic_linkage_kernel_equals_scipy(emul, kind, n, method):
    """lk_generic_kernel (scipy's fast_linkage): rows come out final, in merge order."""
    from scipy.cluster.hierarchy import linkage
    D = _matrix(kind, n, np.random.default_rng(n))
    with np.errstate(all="ignore"):
        ref = linkage(D[np.triu_indices(n, 1)], method)
    assert np.array_equal(_run(emul, method, D, False), ref, equal_nan=True)


@pytest.mark.parametrize("lazy", [0, 1, 2])
@pytest.mark.parametrize("method,kind,n", [("average", "ties", 3), ("average", "blocks", 200), ("average", "jaccard", 131),
                                           ("average", "line", 150), ("complete", "ties", 90), ("weighted", "random", 77),
                                           ("ward", "euclid", 110), ("average", "blocks", 520), ("single", "ties", 140), ("single", "euclid", 65)])
def test_emulated_chain_and_mst_kernels_equal_scipy(emul, method, kind, n, lazy):
    """lk_nn_chain_kernel (0: eager, 1: lazy column updates), lk_nn_chain_list_kernel (2: lazy + live list) and
    lk_mst_single_kernel: rows in merge order, then the stable sort and union-find relabelling scipy applies (the library does
    them on the host: lk_finalize; here the oracle's restatement)."""
    from scipy.cluster.hierarchy import linkage
    from oracle import linkage_oracle as L
    if method == "single" and lazy:
        pytest.skip("the spanning-tree kernel has no lazy variant")
    D = _matrix(kind, n, np.random.default_rng(1000 + n))
    raw = _run(emul, method, D, lazy)
    Z = raw[np.argsort(raw[:, 2], kind="mergesort")]
    L.label(Z, n)
    assert np.array_equal(Z, linkage(D[np.triu_indices(n, 1)], method))


# ------------------------------------------ the whole dendrogram step against the UNMODIFIED reference, without a GPU
@pytest.fixture(scope="module")
def linkage_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_linkage.npz"))


def _native_newick(Z, names):
    import ctypes as C
    from gravity2_b200 import _lib
    lib = _lib.load()
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    n = Z.shape[0] + 1
    arr = (C.c_char_p * n)(*[s.encode() for s in names])
    ln = C.c_size_t(0)
    assert lib.gv2_newick_from_linkage(Z.ctypes.data_as(C.c_void_p), n, arr, None, 0, C.byref(ln)) == 0
    buf = C.create_string_buffer(ln.value + 1)
    assert lib.gv2_newick_from_linkage(Z.ctypes.data_as(C.c_void_p), n, arr, buf, ln.value + 1, C.byref(ln)) == 0
    return buf.value.decode()


@pytest.mark.parametrize("method", ["single", "complete", "average", "weighted", "centroid", "median", "ward"])
@pytest.mark.parametrize("tag", ["A", "B", "T"])
def test_oracle_dendrogram_equals_reference_dist_mat_to_tree(linkage_golden, tag, method):
    """oracle.dist_mat_to_tree (what the GPU tests compare DistMat2Tree / bootstrap_newick with) == the Newick string the
    UNMODIFIED app/utils/dist_mat_to_tree.py:5-21 printed, for all seven linkage methods, plain and log-scaled
    (oracle/make_golden.py linkage)."""
    from oracle import gravity_oracle as O
    D = linkage_golden[f"{tag}_D"]
    names = [f"v{i}" for i in range(len(D))]
    for logscale in (False, True):
        with np.errstate(all="ignore"):
            assert O.dist_mat_to_tree(D.copy(), names, method, logscale) == str(linkage_golden[f"{tag}_{method}_{int(logscale)}"])


@pytest.mark.parametrize("method", ["single", "complete", "average", "weighted", "centroid", "median", "ward"])
@pytest.mark.parametrize("tag", ["A", "T"])
def test_emulated_kernel_and_native_printer_equal_reference_newick(emul, linkage_golden, tag, method):
    """The device kernel's source run on host threads (the default variant of each method), the host-side sort + relabelling
    scipy applies to chain / spanning-tree rows, and the library's own Newick printer (gv2_newick_from_linkage: host code of
    libgravity_b200.so) against the string the unmodified reference printed."""
    from oracle import linkage_oracle as L
    D = linkage_golden[f"{tag}_D"]
    n = len(D)
    names = [f"v{i}" for i in range(n)]
    for logscale in (False, True):
        M = 10 ** D if logscale else D
        raw = _run(emul, method, M, 0 if method in ("single", "centroid", "median") else 2)
        if method in ("centroid", "median"):
            Z = raw
        else:
            Z = raw[np.argsort(raw[:, 2], kind="mergesort")]
            L.label(Z, n)
        assert _native_newick(Z, names) == str(linkage_golden[f"{tag}_{method}_{int(logscale)}"])
