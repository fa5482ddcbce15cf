"""The GOM scoring kernels, run on the CPU: tests/emul/gom_emul.cpp includes the SAME
source the GPU build compiles (gravity2_b200/csrc/gom_score_kernels.cuh) and follows gom_db_build / gom_db_score step by step
as gv2_gom_table_host drives them -- genome CSR, per-GOM point sets and distance caches, pair records, the location-ordered
prefix sums, weighted centroids and aggregates, the light and heavy score kernels with replicates on the lanes.  Checked
against the reference's own known answer for dcor (app/utils/dcor.py:38-41), its goldens and, with bootstrap multiplicities,
the oracle on the resampled table ``loc[:, idx]`` -- 1e-9 relative, as for the GPU tests; exactly degenerate inputs give
exactly 0."""
import os
import subprocess
import warnings

import numpy as np
import pytest

from oracle import gravity_oracle as O
from gravity2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(emul_build):
    return emul_build("gom_emul")


def gom_table(emul, loc, rows, offs, weights=None, all_columns=False):
    """Engine.gom_table on the emulated kernels: (N, G), or (B, N, G) with multiplicities."""
    loc = np.ascontiguousarray(loc, dtype=np.float64)
    n, p = loc.shape
    g = len(offs) - 1
    nb = 1 if weights is None else len(weights)
    blob = loc.tobytes() + np.asarray(offs, dtype=np.int32).tobytes() + np.ascontiguousarray(rows, dtype=np.float64).tobytes()
    if weights is not None:
        blob += np.ascontiguousarray(weights, dtype=np.int32).tobytes()
    out = subprocess.run([emul, str(n), str(p), str(g), str(nb), "1" if all_columns else "0", "0" if weights is None else "1"],
                         input=blob, capture_output=True, timeout=1800)
    assert out.returncode == 0, (out.returncode, out.stderr)
    res = np.frombuffer(out.stdout, dtype=np.float64).reshape(nb, n, g)
    return res if weights is not None else res[0]


def rows_of(db, ids, p):
    rows = [np.asarray(db[g], dtype=np.float64).reshape(-1, p) for g in ids]
    offs = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    return np.concatenate(rows, axis=0), offs


def close(a, b, rtol=1e-9, atol=1e-12):
    err = np.abs(a - b)
    assert a.shape == b.shape and (err <= atol + rtol * np.abs(b)).all(), f"worst abs {err.max():.3e}"


def test_emulated_dcor_known_answer(emul):
    """dcor.py:38-41: dcor([1, 2, 3, 4, 5], [1, 2, 9, 4, 4]) = 0.76267624241686649 -- one GOM whose single member is X, one
    genome whose location vector is Y, every column relevant (the drop-in's dcor)."""
    X = np.array([[1.0, 2, 3, 4, 5]])
    Y = np.array([[1.0, 2, 9, 4, 4]])
    v = gom_table(emul, Y, X, [0, 1], all_columns=True)[0, 0]
    assert abs(v - 0.7626762424168665) <= 1e-12
    # exactly degenerate inputs: the reference returns exactly 0.0 (dcor.py:54-60)
    assert gom_table(emul, np.full((1, 5), 3.0), X, [0, 1], all_columns=True)[0, 0] == 0.0
    assert gom_table(emul, Y, np.full((2, 5), 7.0), [0, 2], all_columns=True)[0, 0] == 0.0


@pytest.mark.parametrize("tag", ["A", "B"])
def test_emulated_gom_table_equals_reference_golden(emul, tag):
    """GOMSignatureTable_Constructor of the unmodified reference on the golden fixtures (reference_golden.npz: `gom`)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    loc, taxo, ids = g[f"{tag}_loc"], g[f"{tag}_taxo"], [str(x) for x in g[f"{tag}_gom_ids"]]
    rows, offs = rows_of(O.gomdb(taxo, loc, ids), ids, loc.shape[1])
    close(gom_table(emul, loc, rows, offs), g[f"{tag}_gom"])


def test_emulated_location_dcor_matrix_equals_reference_golden(emul):
    """Scheme L (a5): one single-member GOM per genome; the reference's cleaned L matrix is the golden S_L."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    loc = g["A_loc"]
    L = gom_table(emul, loc, loc, np.arange(len(loc) + 1))
    ref = g["A_S_L_p1"]
    Lc = np.where(np.isnan(L) | (L < 0), 0.0, L)
    close(Lc, ref, rtol=1e-9, atol=1e-9)


def test_emulated_gom_tables_of_bootstrap_replicates_vs_oracle(emul):
    """Replicates on the lanes (40: two chunks, one sweep of the distance caches) with bootstrap multiplicities against the
    oracle run on the resampled tables loc[:, idx] with the GOM database rebuilt per replicate, as the PL1 loop does
    (pl1_graphs.py:81-100); light and heavy pairs, an empty genome, a genome whose drawn hits may all vanish."""
    t = synth.make_tables(36, 260, 4, 0.16, seed=3)
    loc = t.loc.copy()
    loc[5] = 0.0
    rng = np.random.default_rng(9)
    nb = 40
    idx = [rng.integers(0, 260, 260) for _ in range(nb)]
    w = np.stack([np.bincount(i, minlength=260) for i in idx]).astype(np.int32)
    rows, offs = rows_of(O.gomdb(t.taxo, loc, t.gom_ids), t.gom_ids, 260)
    got = gom_table(emul, loc, rows, offs, weights=w)
    assert got.shape == (nb, 36, 4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for b in (0, 7, 31, 32, 39):
            lb = loc[:, idx[b]]
            ref = O.gom_table(lb, O.gomdb(t.taxo, lb, t.gom_ids), t.gom_ids)
            close(got[b], ref)
        close(gom_table(emul, loc, rows, offs), O.gom_table(loc, O.gomdb(t.taxo, loc, t.gom_ids), t.gom_ids))
    assert np.all(got[:, 5, :] == 0.0)


def test_emulated_gom_heavy_pairs_vs_oracle(emul):
    """Dense location rows: a genome shares ~80 PPHMMs with its own taxon, far above the light kernel's 32-entry records, so
    these pairs go through gom_score_heavy_kernel (rows of M by asynchronous copies, four warps per pair)."""
    t = synth.make_tables(20, 300, 3, 0.3, seed=8)
    assert (t.loc != 0).sum(1).max() > 60
    rng = np.random.default_rng(4)
    idx = [rng.integers(0, 300, 300) for _ in range(3)]
    w = np.stack([np.bincount(i, minlength=300) for i in idx]).astype(np.int32)
    rows, offs = rows_of(O.gomdb(t.taxo, t.loc, t.gom_ids), t.gom_ids, 300)
    got = gom_table(emul, t.loc, rows, offs, weights=w)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for b in range(3):
            lb = t.loc[:, idx[b]]
            close(got[b], O.gom_table(lb, O.gomdb(t.taxo, lb, t.gom_ids), t.gom_ids))
