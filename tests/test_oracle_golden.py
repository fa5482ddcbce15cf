"""Pins the CPU oracle (oracle/gravity_oracle.py) to the reference: the reference's own
known answers and the golden outputs of the unmodified reference functions
(tests/golden/reference_golden.npz, written by oracle/make_golden.py)."""
import warnings

import numpy as np
import pytest

from oracle import gravity_oracle as O

TAGS = ["A", "B"]


def test_dcor_docstring_kat(golden):
    # app/utils/dcor.py:38-41
    v = O.dcor(np.array([[1.], [2.], [3.], [4.], [5.]]), np.array([[1.], [2.], [9.], [4.], [4.]]))
    assert v == pytest.approx(0.76267624241686649, rel=1e-15)
    assert v == float(golden["dcor_kat"])


def test_cent_dist_literals(golden):
    # test/utils/test_dcor.py:29-43
    got = O.cent_dist(np.array([[1, 2], [3, 4], [5, 6]]))
    M = np.array([[0.0, 2.82842712, 5.65685425], [2.82842712, 0.0, 2.82842712], [5.65685425, 2.82842712, 0.0]])
    exp = M - M.mean(1)[:, None] - M.mean(0)[None, :] + M.mean()
    assert np.allclose(got, exp)
    assert np.array_equal(got, golden["cent_dist_123456"])


def test_gomdb_fixture():
    # test/utils/test_gomdb_constructor.py:5-31
    taxo = np.array(["A", "B", "A", "C"])
    loc = np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1], [0, 0, 1]])
    db = O.gomdb(taxo, loc, ["A", "B", "C"])
    assert np.array_equal(db["A"], [[1, 0, 1], [1, 1, 1]])
    assert np.array_equal(db["B"], [[0, 1, 0]])
    assert np.array_equal(db["C"], [[0, 0, 1]])


def test_generate_gom_sigs_fixture(golden):
    # test/utils/test_parallel_gom_sig_generator.py:51-77 (real output is [1, 1, 0])
    loc = np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1]], dtype=float)
    db = {"GOM1": np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1]], dtype=float),
          "GOM2": np.array([[0, 1, 0], [1, 0, 1], [0, 0, 0]], dtype=float)}
    for g in db:
        got = np.array(O.generate_gom_sigs(db[g], loc))
        assert np.array_equal(got, golden[f"gomfix_{g}"])
    assert np.allclose(golden["gomfix_GOM1"], [1.0, 1.0, 0.0])


def test_newick_golden(golden):
    # test/utils/test_dist_mat_to_tree.py:5-19
    D = np.array([[0.0, 0.2, 0.4], [0.2, 0.0, 0.6], [0.4, 0.6, 0.0]])
    s = O.dist_mat_to_tree(D, ["A", "B", "C"], "single", False)
    assert s == "((B:0.200000,A:0.200000):0.200000,C:0.400000);"
    assert s == str(golden["newick_3x3_single"])


@pytest.mark.parametrize("tag", TAGS)
def test_gom_table_matches_reference(golden, tag):
    loc, taxo, ids = golden[f"{tag}_loc"], golden[f"{tag}_taxo"], list(golden[f"{tag}_gom_ids"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = O.gom_table(loc, O.gomdb(taxo, loc, ids), ids)
    assert np.array_equal(got, golden[f"{tag}_gom"], equal_nan=True)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("scheme,p", [("P", 1), ("G", 1), ("L", 1), ("PG", 1), ("PG", 2), ("PL", 1),
                                      ("R", 1), ("RG", 1), ("PR", 1)])
def test_similarity_mat_matches_reference(golden, tag, scheme, p):
    sig, loc, gom, spr = (golden[f"{tag}_{k}"] for k in ("sig", "loc", "gom", "spr"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = O.similarity_mat(sig.copy(), gom, loc, spr.copy(), 0.0, 0, scheme, float(p))
    assert np.array_equal(got, golden[f"{tag}_S_{scheme}_p{p}"])
    if "L" not in scheme:
        fast = O.similarity_mat_fast(sig, gom, spr, 0, scheme, float(p))
        assert np.allclose(fast, got, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("tag", TAGS)
def test_threshold_mutates_caller_table(golden, tag):
    sig, loc, gom = (golden[f"{tag}_{k}"] for k in ("sig", "loc", "gom"))
    s = sig.copy()
    got = O.similarity_mat(s, gom, loc, None, 0.0, 60, "PG", 1.0)
    assert np.array_equal(got, golden[f"{tag}_S_PG_thr60"])
    assert np.array_equal(s, golden[f"{tag}_sig_after_thr60"])


@pytest.mark.parametrize("tag", TAGS)
def test_presence_matrices_match_reference(golden, tag):
    sig = golden[f"{tag}_sig"]
    assert np.array_equal(np.array(O.shared_pphmm_ratio(sig)), golden[f"{tag}_shared"])
    keep = golden[f"{tag}_shared_norm_keep"]
    assert np.array_equal(O.shared_norm_pphmm_ratio(sig[keep]), golden[f"{tag}_shared_norm"])
    if len(keep) < len(sig):
        with pytest.raises(ZeroDivisionError):
            O.shared_norm_pphmm_ratio(sig)


@pytest.mark.parametrize("tag", TAGS)
def test_bootstrap_loops_match_reference(golden, tag):
    sig, loc, taxo = golden[f"{tag}_sig"], golden[f"{tag}_loc"], golden[f"{tag}_taxo"]
    ids = list(golden[f"{tag}_gom_ids"])
    p = sig.shape[1]
    np.random.seed(int(golden[f"{tag}_boot_seed"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r in range(3):
            idx = O.bootstrap_indices(p, sort=False)
            assert idx.dtype == np.int64 and np.array_equal(idx, golden[f"{tag}_pl1_idx"][r])
            D = O.pl1_bootstrap_replicate(sig, taxo, ids, idx)
            assert np.array_equal(D, golden[f"{tag}_pl1_D"][r])
        n_ref = int(golden[f"{tag}_pl2_nref"])
        for r in range(3):
            idx = O.bootstrap_indices(p, sort=True)
            assert np.array_equal(idx, golden[f"{tag}_pl2_idx"][r])
            S, _ = O.pl2_bootstrap_replicate(sig, loc, n_ref, taxo[:n_ref], ids, idx)
            assert np.array_equal(S, golden[f"{tag}_pl2_S"][r])


@pytest.mark.parametrize("tag", TAGS)
def test_weights_equal_resample(golden, tag):
    """Column resample == integer column multiplicities, for GJ and dCor."""
    sig, loc = golden[f"{tag}_sig"], golden[f"{tag}_loc"]
    idx = golden[f"{tag}_pl1_idx"][0]
    w = O.weights_from_indices(idx, sig.shape[1])
    a = O.weighted_gj(sig[:, idx])
    b = O.weighted_gj(sig, w)
    assert np.allclose(a, b, rtol=1e-13, atol=0, equal_nan=True)
    X, Y = loc[:4].T, loc[5].reshape(-1, 1)
    assert O.weighted_dcor(X, Y, w) == pytest.approx(O.dcor(X[idx], Y[idx]), rel=1e-12)


@pytest.mark.parametrize("tag", TAGS)
def test_newick_average_matches_reference(golden, tag):
    S = golden[f"{tag}_S_PG_p1"]
    D = O.dist_from_sim(S.copy())
    n = len(D)
    assert O.dist_mat_to_tree(D, [f"v{i}" for i in range(n)], "average") == str(golden[f"{tag}_newick_avg"])


# ------------------------------------------------- PphmmNeighbourhoodWeight != 0
@pytest.mark.parametrize("tag", ["C", "D"])
@pytest.mark.parametrize("case", [0, 1, 2])
def test_neighbourhood_mode_oracle_is_the_reference(golden_nbhd, tag, case):
    g = golden_nbhd
    w, thr, pw = g[f"{tag}_case{case}"]
    scheme = str(g[f"{tag}_case{case}_scheme"])
    nb = O.neighbourhood_index_lists(g[f"{tag}_naive_loc"], g[f"{tag}_sig"])
    assert nb[0], "fixture must contain neighbourhoods"
    sig = g[f"{tag}_sig"].copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S = O.similarity_mat(sig, g[f"{tag}_gom"], None, None, w, thr, scheme, pw, neighbourhoods=nb)
        S2 = O.similarity_mat(sig, g[f"{tag}_gom"], None, None, w, thr, scheme, pw, neighbourhoods=nb)
    assert np.array_equal(S, g[f"{tag}_case{case}_S"])
    assert np.array_equal(S2, g[f"{tag}_case{case}_S_second_call"])           # starts from the mutated table
    assert np.array_equal(sig, g[f"{tag}_case{case}_sig_after_two_calls"])
    assert not np.array_equal(S, S2)


@pytest.mark.parametrize("tag", ["C", "D"])
def test_neighbourhood_host_logic(golden_nbhd, tag):
    """gravity2_b200.neighbourhood (array formulation) == the loop restatement == the reference."""
    from gravity2_b200 import neighbourhood as NB
    g = golden_nbhd
    neg, pos = O.neighbourhood_index_lists(g[f"{tag}_naive_loc"], g[f"{tag}_sig"])
    ref = np.zeros(g[f"{tag}_sig"].shape, dtype=np.uint8)
    for r, cols in neg.items():
        ref[r, cols] |= NB.NEG
    for r, cols in pos.items():
        ref[r, cols] |= NB.POS
    cls = NB.neighbourhood_classes(g[f"{tag}_naive_loc"], g[f"{tag}_sig"])
    assert np.array_equal(cls, ref)
    for case in range(3):
        w, thr, _ = g[f"{tag}_case{case}"]
        sig = g[f"{tag}_sig"].copy()
        n = len(sig)
        NB.apply_in_place(sig, cls, w, thr, n + 1)
        NB.apply_in_place(sig, cls, w, thr, n + 1)
        assert np.array_equal(sig, g[f"{tag}_case{case}_sig_after_two_calls"])


def test_neighbourhood_host_logic_random():
    from gravity2_b200 import neighbourhood as NB
    rng = np.random.default_rng(7)
    for _ in range(120):
        n, p = int(rng.integers(1, 8)), int(rng.integers(1, 30))
        sig = np.where(rng.random((n, p)) < rng.random(), np.round(rng.gamma(2, 60, (n, p)), 1), 0.0)
        loc = np.where(sig != 0, rng.integers(1, 5, (n, p)) * 1.0, 0.0)
        if rng.random() < 0.3:
            loc[:, rng.integers(0, p)] = 0          # never-hit column: NaN mean, sorts last
        neg, pos = O.neighbourhood_index_lists(loc, sig)
        ref = np.zeros((n, p), dtype=np.uint8)
        for r, cols in neg.items():
            ref[r, cols] |= NB.NEG
        for r, cols in pos.items():
            ref[r, cols] |= NB.POS
        assert np.array_equal(NB.neighbourhood_classes(loc, sig), ref)


# ------------------------------------------------- scipy's linkage, restated (third-party arithmetic on the tree step)
@pytest.mark.parametrize("method", ["average", "complete", "weighted", "ward", "single", "centroid", "median"])
def test_linkage_restatement_is_scipy(method):
    """oracle/linkage_oracle.py (nn_chain / mst_single_linkage / fast_linkage + Heap / label of scipy/cluster/_hierarchy.pyx,
    loop for loop) == the installed scipy, bit for bit, on matrices with and without exact ties."""
    from scipy.cluster.hierarchy import linkage
    from oracle import linkage_oracle as L
    rng = np.random.default_rng(5)
    for trial in range(8):
        n = int(rng.integers(2, 50))
        D = rng.integers(1, 6, (n, n)) / 5.0 if trial % 2 else rng.random((n, n))
        D = np.triu(D, 1)
        D = D + D.T
        dl = D[np.triu_indices_from(D, k=1)]
        with np.errstate(all="ignore"):
            assert np.array_equal(L.linkage(dl, method), linkage(dl, method=method), equal_nan=True)


# ------------------------------------------------------------------ a16: sort_pphmm binary Jaccard (reference golden)
@pytest.mark.parametrize("tag", ["E", "F"])
def test_binary_gj_equals_reference_sort_pphmm(tag):
    """oracle.binary_gj restates app/src/ref_virus_annotator.py:148-158; the golden is what the UNMODIFIED
    RefVirusAnnotator.sort_pphmm handed to DistMat2Tree (oracle/make_golden.py sort_pphmm)."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_sort_pphmm.npz")
    g = dict(np.load(path, allow_pickle=False))
    D = 1 - O.binary_gj(g[f"{tag}_sig"].T)
    ref = g[f"{tag}_traits_D"]
    assert D.shape == ref.shape
    assert np.array_equal(np.isnan(D), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(D, nan=-1.0), np.nan_to_num(ref, nan=-1.0))      # bit for bit
    if tag == "F":
        empty = (g["F_sig"] != 0).sum(axis=0) == 0                                        # empty vs empty column: 0/0
        assert empty[5] and empty[20] and np.array_equal(np.isnan(ref), np.outer(empty, empty))
        assert ref[5, 0] == 1.0                                                           # empty vs non-empty: 0/n = 0
    else:
        leaves = list(range(D.shape[0]))
        assert O.dist_mat_to_tree(D.copy(), leaves, "average", False) == str(g["E_traits_newick"])
