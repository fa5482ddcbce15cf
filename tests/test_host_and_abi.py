"""CPU-only tests: the C-ABI library loads and exports every symbol include/gravity2_b200.h
declares, the ctypes prototypes cover the header, the product path fails loudly without a GPU,
and the host-side logic (GOM membership, resample weights, tile maths) is right."""
import ctypes
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(REPO, "include", "gravity2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gv2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gravity2_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names
    assert lib.gv2_version() >= 100


def test_header_is_plain_c_and_cites_the_reference():
    """The boundary is a C ABI: the header compiles as C99 on its own (no C++ types, no torch types in a signature), and every
    entry point's comment cites the reference interface it replaces (file:line under app/)."""
    import shutil
    import subprocess
    path = os.path.join(REPO, "include", "gravity2_b200.h")
    if shutil.which("gcc") is not None:
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", path],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    src = open(path).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)                    # signatures only, comments stripped
    assert 'extern "C"' in code and "torch" not in code and "std::" not in code and "&" not in code.replace("&&", "")
    assert len(re.findall(r"[a-z_0-9]+\.py:\d+", src)) >= 20


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gravity2_b200
    with pytest.raises(gravity2_b200.engine.GravityNativeError) as ei:
        gravity2_b200.Engine(0)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(gravity2_b200.engine.GravityNativeError):
        gravity2_b200.SimilarityMat_Constructor(np.ones((2, 3)), np.ones((2, 2)), None, None, 0.0, 0, "PG", 1.0)


def test_product_code_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "gravity2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(root, f)).read().replace("oracle/philox_oracle.py", ""), f


def test_num_tiles_and_membership():
    from gravity2_b200 import _lib, gom_membership, weights_from_indices
    lib = _lib.load()
    assert [lib.gv2_num_tiles(n) for n in (0, 1, 128, 129, 1000, 10000)] == [0, 1, 1, 3, 36, 3160]
    taxo = np.array(["A", "B", "A", "C", "Z"])
    assert list(gom_membership(taxo, ["A", "B", "C"])) == [0, 1, 0, 2, -1]
    assert list(gom_membership(taxo[:3], ["A", "B", "C"], n_rows=5)) == [0, 1, 0, -1, -1]
    w = weights_from_indices([0, 0, 3, 2, 3], 5)
    assert w.dtype == np.int32 and list(w) == [2, 0, 1, 2, 0]


def test_reference_index_stream_is_randint():
    """np.random.choice(list(range(P)), P, replace=True) == legacy randint stream."""
    from gravity2_b200 import draw_resample_indices
    np.random.seed(5)
    a = draw_resample_indices(1000)
    np.random.seed(5)
    b = np.random.randint(0, 1000, size=1000)
    assert a.dtype == np.int64 and np.array_equal(a, b)
    np.random.seed(5)
    c = draw_resample_indices(1000, sort=True)
    assert np.array_equal(c, np.sort(b))


def test_gomdb_constructor_fixture():
    from gravity2_b200 import GOMDB_Constructor
    taxo = np.array(["A", "B", "A", "C"])
    loc = np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1], [0, 0, 1]])
    db = GOMDB_Constructor(taxo, loc, ["A", "B", "C"])
    assert np.array_equal(db["A"], [[1, 0, 1], [1, 1, 1]]) and np.array_equal(db["C"], [[0, 0, 1]])


def test_philox_oracle_known_answers():
    from oracle.philox_oracle import philox4x32_10
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)


def test_native_newick_matches_the_reference_printer(golden):
    """gv2_newick_from_linkage (host only: no GPU needed) == to_tree + GetNewick (app/utils/get_newick.py:1-13) on scipy
    linkage matrices, including the reference's own golden (test/utils/test_dist_mat_to_tree.py:17 uses method single;
    the printer does not care) and the reference-produced Newick of the synthetic fixtures."""
    import ctypes as C
    from scipy.cluster.hierarchy import linkage
    from gravity2_b200 import _lib
    from oracle import gravity_oracle as O
    lib = _lib.load()

    def native(Z, names):
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        n = Z.shape[0] + 1
        arr = (C.c_char_p * n)(*[s.encode() for s in names])
        ln = C.c_size_t(0)
        assert lib.gv2_newick_from_linkage(Z.ctypes.data_as(C.c_void_p), n, arr, None, 0, C.byref(ln)) == 0
        buf = C.create_string_buffer(ln.value + 1)
        assert lib.gv2_newick_from_linkage(Z.ctypes.data_as(C.c_void_p), n, arr, buf, ln.value + 1, C.byref(ln)) == 0
        return buf.value.decode()

    D3 = np.array([[0.0, 0.2, 0.4], [0.2, 0.0, 0.6], [0.4, 0.6, 0.0]])
    Z = linkage(D3[np.triu_indices_from(D3, k=1)], method="single")
    assert native(Z, ["A", "B", "C"]) == str(golden["newick_3x3_single"])
    rng = np.random.default_rng(3)
    for n in (2, 3, 7, 40, 300):
        X = rng.random((n, 4))
        D = np.round(np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1)), 2)      # rounded: plenty of exact ties
        names = [f"v{i}|x" for i in range(n)]
        Z = linkage(D[np.triu_indices_from(D, k=1)], method="average")
        assert native(Z, names) == O.dist_mat_to_tree(D.copy(), names, "average", False)
    for tag in ("A", "B"):
        S = golden[f"{tag}_S_PG_p1"]
        D0 = 1 - S
        D0[D0 < 0] = 0
        names = [f"v{i}" for i in range(len(S))]
        Z = linkage(D0[np.triu_indices_from(D0, k=1)], method="average")
        assert native(Z, names) == str(golden[f"{tag}_newick_avg"])
