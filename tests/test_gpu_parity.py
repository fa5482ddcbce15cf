"""GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI
(libgravity_b200.so) and is compared with the CPU oracle / the reference's golden outputs.
Tolerances: bit-exact for counts and indices; rel 1e-6 on score-based similarities
(BASELINE.json north_star); the fp64 GOM/dCor path is held to 1e-9."""
import os
import tempfile
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import gravity_oracle as O
from oracle import philox_oracle
from gravity2_b200 import synth

RTOL = 1e-6        # north_star tolerance for score-based similarities
ATOL = 1e-12       # entries below this are compared absolutely


@pytest.fixture(scope="module")
def eng():
    import gravity2_b200
    return gravity2_b200.default_engine()


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype == np.float64
    err = np.abs(a - b)
    ok = err <= atol + rtol * np.abs(b)
    assert ok.all(), f"max rel err {np.max(err / np.maximum(np.abs(b), 1e-300)):.3e}, worst abs {err.max():.3e}"


# ------------------------------------------------------------------ golden fixtures (reference outputs)
@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("scheme,p", [("P", 1), ("G", 1), ("L", 1), ("PG", 1), ("PG", 2), ("PL", 1), ("R", 1), ("RG", 1), ("PR", 1)])
def test_similarity_golden(golden, eng, tag, scheme, p):
    import gravity2_b200 as G
    sig, loc, gom, spr = (golden[f"{tag}_{k}"].copy() for k in ("sig", "loc", "gom", "spr"))
    S = G.SimilarityMat_Constructor(sig, gom, loc, spr, 0.0, 0, scheme, float(p), fnames=None)
    close(S, golden[f"{tag}_S_{scheme}_p{p}"])
    assert np.array_equal(S, S.T)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_threshold_golden(golden, eng, tag):
    import gravity2_b200 as G
    sig = golden[f"{tag}_sig"].copy()
    S = G.SimilarityMat_Constructor(sig, golden[f"{tag}_gom"], golden[f"{tag}_loc"], None, 0.0, 60, "PG", 1.0)
    close(S, golden[f"{tag}_S_PG_thr60"])
    assert np.array_equal(sig, golden[f"{tag}_sig_after_thr60"])      # caller's table mutated like the reference


@pytest.mark.parametrize("tag", ["C", "D"])
@pytest.mark.parametrize("case", [0, 1, 2])
def test_neighbourhood_weight_golden(golden_nbhd, eng, tag, case):
    """PphmmNeighbourhoodWeight != 0: the reference's visit-count dependent in-place row scaling (quirk Q1),
    two consecutive calls on the same (mutated) table, against the unmodified reference's outputs."""
    import gravity2_b200 as G
    from conftest import write_reference_csvs
    g = golden_nbhd
    w, thr, pw = g[f"{tag}_case{case}"]
    scheme = str(g[f"{tag}_case{case}_scheme"])
    sig = g[f"{tag}_sig"].copy()
    with tempfile.TemporaryDirectory() as tmp:
        fn = write_reference_csvs(tmp, g[f"{tag}_sig"], g[f"{tag}_naive_loc"])
        S = G.SimilarityMat_Constructor(sig, g[f"{tag}_gom"], None, None, w, thr, scheme, pw, fn)
        S2 = G.SimilarityMat_Constructor(sig, g[f"{tag}_gom"], None, None, w, thr, scheme, pw, fn)
    close(S, g[f"{tag}_case{case}_S"])
    close(S2, g[f"{tag}_case{case}_S_second_call"])
    assert np.array_equal(S, S.T)
    assert np.array_equal(sig, g[f"{tag}_case{case}_sig_after_two_calls"])    # caller's table left as the reference leaves it


@pytest.mark.parametrize("scheme", ["P", "PG", "PR", "PL"])
def test_neighbourhood_weight_vs_oracle(eng, scheme):
    """Larger seeded table, every scheme with a P component, against the loop restatement."""
    import gravity2_b200 as G
    from conftest import write_reference_csvs
    t = synth.make_tables(45, 160, 4, 0.45, seed=77)
    rng = np.random.default_rng(5)
    naive = np.where(t.sig != 0, np.arange(160)[None, :] * 50.0 + rng.integers(1, 40, t.sig.shape), 0.0)
    spr = rng.random((45, 45)); spr = (spr + spr.T) / 2
    db = O.gomdb(t.taxo, t.loc, t.gom_ids)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gom = O.gom_table(t.loc, db, t.gom_ids)
        nb = O.neighbourhood_index_lists(naive, t.sig)
        assert len(nb[0]) > 10
        s_ref = t.sig.copy()
        ref = O.similarity_mat(s_ref, gom, t.loc, spr.copy(), 0.1, 25, scheme, 1.5, neighbourhoods=nb)
    s_dev = t.sig.copy()
    with tempfile.TemporaryDirectory() as tmp:
        fn = write_reference_csvs(tmp, t.sig, naive)
        S = G.SimilarityMat_Constructor(s_dev, gom, t.loc, spr.copy(), 0.1, 25, scheme, 1.5, fn)
    close(S, ref)
    assert np.array_equal(s_dev, s_ref)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_gom_table_golden(golden, eng, tag):
    import gravity2_b200 as G
    loc, taxo, ids = golden[f"{tag}_loc"], golden[f"{tag}_taxo"], list(golden[f"{tag}_gom_ids"])
    db = G.GOMDB_Constructor(taxo, loc, ids)
    tab = G.GOMSignatureTable_Constructor(loc, db, ids, {"N_CPUs": 1})
    assert tab.shape == golden[f"{tag}_gom"].shape
    close(tab, golden[f"{tag}_gom"], rtol=1e-9)
    tab4 = G.GOMSignatureTable_Constructor(loc, db, ids, 0)           # 4-arg twin (quirk Q10)
    assert np.array_equal(tab, tab4)


def test_dcor_kat_and_fixture(golden, eng):
    import gravity2_b200 as G
    v = G.dcor(np.array([[1.], [2.], [3.], [4.], [5.]]), np.array([[1.], [2.], [9.], [4.], [4.]]))
    assert v == pytest.approx(0.76267624241686649, rel=1e-12)
    loc = np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1]], dtype=float)
    db = {"GOM1": np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1]], dtype=float),
          "GOM2": np.array([[0, 1, 0], [1, 0, 1], [0, 0, 0]], dtype=float)}
    for g in db:
        got = G.generate_gom_sigs(g, loc, db)
        assert isinstance(got, list) and all(isinstance(x, float) for x in got)
        assert np.allclose(got, golden[f"gomfix_{g}"], rtol=1e-12, atol=1e-15)
    # edge cases: n_rel = 0, 1; all-zero GOM; identical vectors
    z = np.zeros((1, 5))
    assert G.generate_gom_sigs("g", z, {"g": np.zeros((2, 5))}) == [0.0]
    one = np.array([[0., 3., 0., 0., 0.]])
    assert G.generate_gom_sigs("g", one, {"g": np.zeros((2, 5))}) == [0.0]
    x = np.array([[5., -2., 7., 0., 1.]])
    assert G.generate_gom_sigs("g", x, {"g": x.copy()})[0] == pytest.approx(1.0, rel=1e-12)
    assert G.generate_gom_sigs("g", x, {"g": np.zeros((0, 5))}) == [0.0]


def test_dcor_multidim_random(eng):
    import gravity2_b200 as G
    rng = np.random.default_rng(5)
    for n, m in [(2, 1), (7, 3), (40, 6), (129, 2)]:
        X = rng.normal(size=(n, m)) * (rng.random((n, m)) < 0.7)
        Y = rng.normal(size=(n, 1)) * (rng.random((n, 1)) < 0.7)
        assert G.dcor(X, Y) == pytest.approx(O.dcor(X, Y), rel=1e-10, abs=1e-14)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_presence_golden(golden, eng, tag):
    import gravity2_b200 as G
    import pandas as pd
    sig = golden[f"{tag}_sig"]
    n, p = sig.shape
    with tempfile.TemporaryDirectory() as tmp:
        df = pd.DataFrame(sig, columns=[f"PPHMM {c}" for c in range(p)])
        df.insert(0, "Virus name", [f"v{i}" for i in range(n)])
        f = os.path.join(tmp, "sigs.csv")
        df.to_csv(f)                                                   # as app/src/pl1_graphs.py:307-315
        fn = {"PphmmAndGomSigs": f}
        got = G.shared_pphmm_ratio(list(range(n)), fn, None)
        assert isinstance(got, list) and isinstance(got[0][0], int)
        assert np.array_equal(np.array(got), golden[f"{tag}_shared"])
        keep = list(golden[f"{tag}_shared_norm_keep"])
        assert np.array_equal(G.shared_norm_pphmm_ratio(keep, fn, None), golden[f"{tag}_shared_norm"])
        if len(keep) < n:
            with pytest.raises(ZeroDivisionError):
                G.shared_norm_pphmm_ratio(list(range(n)), fn, None)
    bj = G.binary_jaccard(sig.T)
    assert np.array_equal(bj, O.binary_gj(sig.T), equal_nan=True)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_bootstrap_golden(golden, eng, tag):
    """PL1 and PL2 loop bodies: reference index stream (bit-exact) and replicate matrices."""
    import gravity2_b200 as G
    sig, loc, taxo = golden[f"{tag}_sig"], golden[f"{tag}_loc"], golden[f"{tag}_taxo"]
    ids = list(golden[f"{tag}_gom_ids"])
    payload = {"N_Bootstrap": 3, "SimilarityMeasurementScheme": "PG", "p": 1.0,
               "PphmmNeighbourhoodWeight": 0.0, "PphmmSigScoreThreshold": 0}
    np.random.seed(int(golden[f"{tag}_boot_seed"]))
    # each yielded matrix is a view into the pinned host ring, valid until the next is requested: copy to keep
    got = [m.copy() for m in G.bootstrap_distance_matrices(sig, loc, taxo, ids, payload, pipeline="pl1")]
    assert len(got) == 3
    for r in range(3):
        # D = 1 - S: the 1e-6 relative bound on S is an absolute 1e-6 bound on D
        close(got[r], golden[f"{tag}_pl1_D"][r], rtol=0, atol=1e-6)
        assert np.all(np.diag(got[r])[np.diag(golden[f"{tag}_pl1_D"][r]) == 0] == 0)
    n_ref = int(golden[f"{tag}_pl2_nref"])
    got = [m.copy() for m in G.bootstrap_distance_matrices(sig, loc, taxo[:n_ref], ids, payload, pipeline="pl2", distance=False)]
    for r in range(3):
        close(got[r], golden[f"{tag}_pl2_S"][r])
    # the global numpy stream advanced exactly as in the reference run: next draw matches a fresh replay
    nxt = np.random.randint(0, 1 << 30)
    np.random.seed(int(golden[f"{tag}_boot_seed"]))
    for _ in range(3):
        O.bootstrap_indices(sig.shape[1])
    for _ in range(3):
        O.bootstrap_indices(sig.shape[1], sort=True)
    assert nxt == np.random.randint(0, 1 << 30)


# ------------------------------------------------------------------------ seeded synthetic tables vs oracle
@pytest.mark.parametrize("n,p,g,dens", [(1, 1, 1, 1.0), (3, 5, 2, 1.0), (127, 31, 5, 0.5), (129, 33, 5, 0.3),
                                        (300, 2049, 12, 0.05), (260, 4200, 9, 1.0)])
def test_similarity_shapes_vs_oracle(eng, n, p, g, dens):
    """ragged sizes around the 128-genome tile / 32-column block / 1024-column slab edges"""
    t = synth.make_tables(n, p, g, dens, seed=7 + n)
    gom = np.random.default_rng(n).random((n, g)) * (np.random.default_rng(n + 1).random((n, g)) < 0.8)
    for scheme in ("P", "G", "PG"):
        S = eng.similarity(t.sig, gom, None, scheme, 1.0)
        ref = O.similarity_mat_fast(t.sig, gom, None, 0, scheme, 1.0)
        close(S, ref)
        assert np.array_equal(S, S.T)
    D = eng.similarity(t.sig, gom, None, "PG", 1.0, distance=True)
    close(D, O.dist_from_sim(O.similarity_mat_fast(t.sig, gom, None, 0, "PG", 1.0)), rtol=0, atol=1e-6)


def test_similarity_nan_and_zero_rows(eng):
    t = synth.make_tables(40, 200, 4, 0.2, seed=3)
    gom = np.random.default_rng(0).random((40, 4))
    gom[5, 2] = np.nan                     # dcor's sqrt of a negative sum (Q5): the whole row's G similarity -> 0
    t.sig[7] = 0
    gom[9] = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = O.similarity_mat(t.sig.copy(), gom, t.loc, None, 0.0, 0, "PG", 1.0)
    S = eng.similarity(t.sig, gom, None, "PG", 1.0)
    close(S, ref)
    assert S[7, 7] == 0.0 and S[5, 5] == 0.0 and S[9, 9] == 0.0


def test_cfg2_full_size_vs_oracle_and_dendrogram(eng):
    """BASELINE config 2 shape: 1k genomes x 5k PPHMMs, PG, base matrix; plus identical UPGMA merges."""
    t = synth.make_config("cfg2")
    import gravity2_b200 as G
    gor = G.gom_membership(t.taxo, t.gom_ids)
    order = np.argsort(gor, kind="stable")
    offs = np.concatenate([[0], np.cumsum(np.bincount(gor, minlength=len(t.gom_ids)))])
    gom = eng.gom_table(t.loc, t.loc[order], offs)
    # oracle GOM table on a row subsample (the reference's n^2 dcor is slow): 40 genomes x all GOMs
    rows = np.arange(0, 1000, 25)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_gom = O.gom_table(t.loc[rows], O.gomdb(t.taxo, t.loc, t.gom_ids), t.gom_ids)
    close(gom[rows], ref_gom, rtol=1e-9)
    S = eng.similarity(t.sig, gom, None, "PG", 1.0)
    ref = O.similarity_mat_fast(t.sig, gom, None, 0, "PG", 1.0)
    close(S, ref)
    assert np.all(np.diag(S) == 1.0)
    Z1 = O.linkage_merges(O.dist_from_sim(S.copy()))
    Z0 = O.linkage_merges(O.dist_from_sim(ref.copy()))
    same = np.array_equal(Z1[:, :2], Z0[:, :2])
    if not same:   # tie-aware: merges may swap only where heights tie within 1e-9
        bad = np.where((Z1[:, :2] != Z0[:, :2]).any(axis=1))[0]
        assert np.allclose(Z1[bad, 2], Z0[bad, 2], rtol=0, atol=1e-7), "dendrogram topology differs beyond ties"
    assert np.allclose(Z1[:, 2], Z0[:, 2], rtol=0, atol=2e-6)


def test_weighted_gom_and_bootstrap_vs_resampled_oracle(eng):
    """column multiplicities == literal column resampling (PL2 body), N=60, P=400"""
    import gravity2_b200 as G
    t = synth.make_tables(60, 400, 5, 0.08, seed=11)
    n_ref = 45
    np.random.seed(99)
    idx = [O.bootstrap_indices(400, sort=True) for _ in range(4)]
    w = np.stack([O.weights_from_indices(i, 400) for i in idx])
    gor = G.gom_membership(t.taxo[:n_ref], t.gom_ids)
    S = eng.bootstrap(t.sig, t.loc, gor, len(t.gom_ids), w, "PG", 1.0, distance=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r in range(4):
            ref, _ = O.pl2_bootstrap_replicate(t.sig, t.loc, n_ref, t.taxo[:n_ref], t.gom_ids, idx[r])
            close(S[r], ref)
    # all-ones weights reproduce the base matrix
    ones = np.ones((1, 400), dtype=np.int32)
    S1 = eng.bootstrap(t.sig, t.loc, gor, len(t.gom_ids), ones, "PG", 1.0, distance=False)[0]
    db = O.gomdb(t.taxo[:n_ref], t.loc[:n_ref], t.gom_ids)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        base = O.similarity_mat(t.sig.copy(), O.gom_table(t.loc, db, t.gom_ids), t.loc, None, 0.0, 0, "PG", 1.0)
    close(S1, base)


def test_bootstrap_many_replicates_p_scheme(eng):
    """70 replicates (more than one 32-lane chunk), scheme P, against the weighted oracle"""
    t = synth.make_tables(150, 700, 6, 0.1, seed=21)
    rng = np.random.default_rng(4)
    w = np.stack([np.bincount(rng.integers(0, 700, 700), minlength=700) for _ in range(70)]).astype(np.int32)
    D = eng.bootstrap(t.sig, None, None, 0, w, "P", 1.0, distance=True)
    for r in (0, 31, 32, 69):
        ref = O.dist_from_sim(O.similarity_mat_fast(t.sig, None, None, 0, "P", 1.0, w=w[r]))
        close(D[r], ref, rtol=0, atol=1e-6)


def test_shared_counts_large_bit_exact(eng):
    rng = np.random.default_rng(8)
    for n, p in [(1, 1), (130, 65), (257, 3000), (300, 1025)]:
        tab = rng.random((n, p)) * (rng.random((n, p)) < 0.1)
        inter, nnz = eng.shared_counts(tab)
        nz = (tab != 0).astype(np.int32)
        assert np.array_equal(inter, nz @ nz.T)
        assert np.array_equal(nnz, nz.sum(1))


def test_philox_indices_bit_exact(eng):
    for seed, nb, p in [(0, 2, 7), (0xDEADBEEFCAFEF00D, 3, 1001)]:
        got = eng.philox_indices(seed, nb, p)
        assert got.dtype == np.int64
        assert np.array_equal(got, philox_oracle.philox_indices(seed, nb, p))


def test_streaming_job_matches_batch_and_oracle(eng):
    """Job: tables uploaded once, 150 PG replicates streamed through the rings (tensor-core blocks of
    128 + a remainder, ring ping-pong, callback hand-over) == the all-at-once entry point == oracle."""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    t = synth.make_tables(140, 900, 6, 0.06, seed=31)
    rng = np.random.default_rng(9)
    w = np.stack([np.bincount(rng.integers(0, 900, 900), minlength=900) for _ in range(150)]).astype(np.int32)
    gor = G.gom_membership(t.taxo, t.gom_ids)
    job = Job(eng, t.sig, t.sig, gor, len(t.gom_ids), "PG", 1.0, distance=True, ring_bytes=7 * 140 * 140 * 8)
    seen = {}
    job.for_each_replicate(w, lambda b, m: seen.__setitem__(b, m.copy()))
    assert sorted(seen) == list(range(150))
    gen = [m.copy() for m in job.replicates(w[:5])]
    base = job.base()
    job.close()
    allatonce = eng.bootstrap(t.sig, t.sig, gor, len(t.gom_ids), w, "PG", 1.0, distance=True)
    for b in range(150):
        assert np.array_equal(seen[b], allatonce[b])
    for b in range(5):      # 5 replicates run on the fp32 CUDA-core kernel, 150 on the tensor cores
        close(gen[b], allatonce[b], rtol=0, atol=1e-6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for b in (0, 77, 149):
            idx = np.repeat(np.arange(900), w[b])
            ref = O.pl1_bootstrap_replicate(t.sig, t.taxo, t.gom_ids, idx)
            close(seen[b], ref, rtol=0, atol=1e-6)
        db = O.gomdb(t.taxo, t.sig, t.gom_ids)          # PL1: signature table doubles as location table
        ref0 = O.dist_from_sim(O.similarity_mat(t.sig.copy(), O.gom_table(t.sig, db, t.gom_ids), t.sig, None, 0.0, 0, "PG", 1.0))
    close(base, ref0, rtol=0, atol=1e-6)


@pytest.mark.parametrize("n,nb", [(140, 23), (2, 3), (1, 2)])
def test_condensed_delivery_equals_squareform(eng, n, nb):
    """Replicates delivered in scipy's condensed form (half the D2H bytes) == squareform of the full matrices, callback
    hand-over and generator, through several ring wraps; the drop-in generator's condensed option; then the reference's own
    tree code on the condensed vector."""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    from scipy.spatial.distance import squareform
    t = synth.make_tables(n, 600, max(1, min(4, n)), 0.08, seed=5 + n)
    rng = np.random.default_rng(n)
    w = np.stack([np.bincount(rng.integers(0, 600, 600), minlength=600) for _ in range(nb)]).astype(np.int32)
    gor = G.gom_membership(t.taxo, t.gom_ids)
    job = Job(eng, t.sig, t.sig, gor, len(t.gom_ids), "PG", 1.0, distance=True, ring_bytes=5 * n * n * 8)
    full, cond = {}, {}
    job.for_each_replicate(w, lambda b, m: full.__setitem__(b, m.copy()))
    job.for_each_replicate(w, lambda b, v: cond.__setitem__(b, v.copy()), condensed=True)
    gen = [v.copy() for v in job.replicates(w[:4], condensed=True)]
    job.close()
    assert sorted(cond) == list(range(nb))
    for b in range(nb):
        assert cond[b].shape == (n * (n - 1) // 2,)
        assert np.array_equal(cond[b], squareform(full[b], checks=False) if n > 1 else np.zeros(0))
    for b in range(min(4, nb)):
        assert np.array_equal(gen[b], cond[b])
    if n > 2:
        from scipy.cluster.hierarchy import linkage
        assert np.array_equal(linkage(cond[0], method="average"), eng.linkage(full[0], "average"))
        payload = {"N_Bootstrap": 3, "SimilarityMeasurementScheme": "PG", "p": 1.0}
        np.random.seed(4)
        a = [m.copy() for m in G.bootstrap_distance_matrices(t.sig, None, t.taxo, t.gom_ids, payload, pipeline="pl1")]
        np.random.seed(4)
        c = [v.copy() for v in G.bootstrap_distance_matrices(t.sig, None, t.taxo, t.gom_ids, payload, pipeline="pl1", condensed=True)]
        assert all(np.array_equal(squareform(x, checks=False), y) for x, y in zip(a, c))


def test_tensor_core_path_is_exact_integer_arithmetic(eng):
    """scheme P replicates through tcgen05 (>= 8 replicates): error is the 2^-33 input quantisation only"""
    t = synth.make_tables(200, 1500, 5, 0.3, seed=41)
    rng = np.random.default_rng(2)
    w = np.stack([np.bincount(rng.integers(0, 1500, 1500), minlength=1500) for _ in range(24)]).astype(np.int32)
    S = eng.bootstrap(t.sig, None, None, 0, w, "P", 1.0, distance=False)
    for b in (0, 23):
        ref = O.similarity_mat_fast(t.sig, None, None, 0, "P", 1.0, w=w[b])
        close(S[b], ref, rtol=2e-8)
        assert np.array_equal(S[b], S[b].T) and np.all(np.diag(S[b]) == 1.0)
    # duplicate genomes: exactly 1.0 off the diagonal (integer sums cancel exactly)
    t.sig[10] = t.sig[3]
    S = eng.bootstrap(t.sig, None, None, 0, w[:8], "P", 1.0, distance=False)
    assert np.all(S[:, 3, 10] == 1.0)


@pytest.mark.parametrize("n,p,dens,nb", [(150, 700, 0.1, 5), (200, 1500, 0.6, 24), (131, 40000, 0.02, 300),
                                         (40, 3000, 1.0, 9), (300, 900, 0.03, 600), (170, 900, 0.05, 3),
                                         (140, 2000, 0.08, 12), (140, 2000, 0.08, 13)])
def test_replicate_paths_agree(eng, monkeypatch, n, p, dens, nb):
    """The three replicate paths (CSR walk of sparse tables, tensor-core contraction, fp32 tile kernel) against the
    weighted oracle and each other.  Rows longer than one 256-entry segment, a table wider than one 32768-column
    window, empty and duplicate genomes, diagonal blocks, all three replicate-lane configurations, and the lanes-are-pairs
    kernel of launches with at most 12 replicates (1, 2, 3 replicate quads per lane) beside the streaming kernel."""
    t = synth.make_tables(n, p, 5, dens, seed=n + nb)
    t.sig[7] = 0.0
    t.sig[n - 1] = t.sig[2]
    rng = np.random.default_rng(nb)
    w = np.stack([np.bincount(rng.integers(0, p, p), minlength=p) for _ in range(nb)]).astype(np.int32)
    out = {}
    for path in ("sparse", "mma", "fp32"):
        monkeypatch.setenv("GV2_BOOT_PATH", path)
        out[path] = eng.bootstrap(t.sig, None, None, 0, w, "P", 1.0, distance=False)
    # the sparse path with its pair lists matched on the fly inside the kernel instead of precomputed: the same integers
    monkeypatch.setenv("GV2_BOOT_PATH", "sparse")
    monkeypatch.setenv("GV2_SPARSE_LISTS", "0")
    on_the_fly = eng.bootstrap(t.sig, None, None, 0, w, "P", 1.0, distance=False)
    monkeypatch.delenv("GV2_SPARSE_LISTS")
    monkeypatch.delenv("GV2_BOOT_PATH")
    if p <= 32768 and dens <= 0.1:
        assert np.array_equal(on_the_fly, out["sparse"])
    else:               # row slices / column windows of the on-the-fly kernel meet in fp64
        assert np.allclose(on_the_fly, out["sparse"], rtol=1e-14, atol=0)
    for b in sorted({0, nb // 2, nb - 1}):
        ref = O.similarity_mat_fast(t.sig, None, None, 0, "P", 1.0, w=w[b])
        close(out["sparse"][b], ref, rtol=1e-12)                         # 48-bit fixed point, exact integer sums
        close(out["mma"][b], ref)                                            # 32-bit fixed point: max(T) * 2^-33 per element
        close(out["fp32"][b], ref)
    # duplicate genomes: the integer sums cancel exactly unless a dense row had to be sliced or the table spans two
    # column windows (partial sums then meet in fp64)
    dup = out["sparse"][:, 2, n - 1]
    assert np.all(dup == 1.0) if dens <= 0.1 and p <= 32768 else np.all(np.abs(dup - 1.0) <= 2e-15)
    bad = np.argwhere(out["sparse"][:, 7, :] != 0.0)
    assert len(bad) == 0, f"empty genome: similarity must be exactly 0; first offenders (replicate, j): {bad[:5].tolist()}"
    for b in (0, nb - 1):
        assert np.array_equal(out["sparse"][b], out["sparse"][b].T)


def test_device_primitives_packed_tiles_roundtrip(eng):
    """C-ABI device-pointer primitives (what the multi-GPU path uses): prepare -> min sums for two tile ranges ->
    packed epilogue -> unpack == the one-shot host entry point."""
    import ctypes as C
    import torch
    from gravity2_b200 import _lib
    lib = _lib.load()
    ctx = C.c_void_p()
    assert lib.gv2_create_on_stream(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(ctx)) == 0
    t = synth.make_tables(300, 777, 4, 0.2, seed=77)
    n, p = t.sig.shape
    n_pad, p_pad = 384, 800
    ntl = int(lib.gv2_num_tiles(n))
    assert ntl == 6
    d_sig = torch.from_numpy(t.sig).cuda()
    tab = torch.empty((n_pad, p_pad), dtype=torch.float32, device="cuda")
    rs = torch.empty(n, dtype=torch.float64, device="cuda")
    nanf = torch.empty(n, dtype=torch.uint8, device="cuda")
    vp = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    assert lib.gv2_dev_prepare_table(ctx, vp(d_sig), n, p, p, 0.0, 0, vp(tab), n_pad, p_pad, vp(rs), vp(nanf)) == 0
    out = torch.full((n, n), -1.0, dtype=torch.float64, device="cuda")
    for (a, b) in [(0, 4), (4, 6)]:
        ws = torch.empty((b - a, 128 * 128), dtype=torch.float64, device="cuda")
        packed = torch.empty((b - a, 128 * 128), dtype=torch.float64, device="cuda")
        assert lib.gv2_dev_gj_min_sums(ctx, vp(tab), n_pad, p_pad, a, b, None, 1, 1, vp(ws)) == 0
        comp = _lib.GjComponent(min_sums=ws.data_ptr(), rowsums=rs.data_ptr(), nanflags=nanf.data_ptr(), ksplit=1)
        assert lib.gv2_dev_gj_epilogue(ctx, n, a, b, 1, C.byref(comp), None, None, _lib.SCHEMES["P"], 1.0,
                                       _lib.OUT_SIMILARITY, 1, vp(packed), 0) == 0, lib.gv2_last_error(ctx)
        assert lib.gv2_dev_unpack_tiles(ctx, n, a, b, vp(packed), vp(out)) == 0
    torch.cuda.synchronize()
    ref = eng.similarity(t.sig, None, None, "P", 1.0)
    assert np.array_equal(out.cpu().numpy(), ref)
    lib.gv2_destroy(ctx)


@pytest.mark.parametrize("scheme,pw,kind", [("PG", 1.0, "D"), ("PG", 2.0, "S"), ("G", 1.0, "S"), ("G", 0.5, "D")])
def test_fused_tail_equals_separate_kernels(eng, scheme, pw, kind):
    """gv2_dev_gj_fused (pair sums in registers + epilogue) against gv2_dev_gj_min_sums + gv2_dev_gj_epilogue on the same
    device tables: two replicate tables of 232 columns (two 128-column slabs in the separate path), a genome with a NaN
    entry in each component, an all-zero genome, a duplicate genome; n not a multiple of 64."""
    import ctypes as C
    import torch
    from gravity2_b200 import _lib
    lib = _lib.load()
    ctx = C.c_void_p()
    assert lib.gv2_create_on_stream(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(ctx)) == 0
    vp = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    rng = np.random.default_rng(11)
    n, p, g, nb = 301, 500, 232, 2
    n_pad, p_pad, g_pad = 384, 512, 256
    ntl = int(lib.gv2_num_tiles(n))
    sig = synth.make_tables(n, p, 4, 0.2, seed=5).sig
    sig[9] = 0.0; sig[40] = sig[3]; sig[17, 5] = np.nan
    gom = rng.random((nb, n, g)) * (rng.random((nb, n, g)) < 0.7)
    gom[:, 9] = 0.0; gom[:, 40] = gom[:, 3]; gom[1, 23, 7] = np.nan
    out_kind = _lib.OUT_DISTANCE if kind == "D" else _lib.OUT_SIMILARITY

    def prep(a64, k, k_pad):
        d = torch.from_numpy(np.ascontiguousarray(a64)).cuda()
        tab = torch.empty((n_pad, k_pad), dtype=torch.float32, device="cuda")
        rs = torch.empty(n, dtype=torch.float64, device="cuda")
        nf = torch.empty(n, dtype=torch.uint8, device="cuda")
        assert lib.gv2_dev_prepare_table(ctx, vp(d), n, k, k, 0.0, 0, vp(tab), n_pad, k_pad, vp(rs), vp(nf)) == 0
        return tab, rs, nf

    ptab, prs, pnf = prep(sig, p, p_pad)
    wsp = torch.empty((ntl, 128 * 128), dtype=torch.float64, device="cuda")
    assert lib.gv2_dev_gj_min_sums(ctx, vp(ptab), n_pad, p_pad, 0, ntl, None, 1, 1, vp(wsp)) == 0
    cp = _lib.GjComponent(min_sums=wsp.data_ptr(), rowsums=prs.data_ptr(), nanflags=pnf.data_ptr(), ksplit=1)
    gtabs = torch.empty((nb, n_pad, g_pad), dtype=torch.float32, device="cuda")
    grs = torch.empty((nb, n), dtype=torch.float64, device="cuda")
    gnf = torch.empty((nb, n), dtype=torch.uint8, device="cuda")
    sep = torch.empty((nb, n, n), dtype=torch.float64, device="cuda")
    for b in range(nb):
        tab, rs, nf = prep(gom[b], g, g_pad)
        gtabs[b], grs[b], gnf[b] = tab, rs, nf
        wsg = torch.empty((ntl, 128 * 128), dtype=torch.float64, device="cuda")
        assert lib.gv2_dev_gj_min_sums(ctx, vp(tab), n_pad, g_pad, 0, ntl, None, 1, 1, vp(wsg)) == 0
        cg = _lib.GjComponent(min_sums=wsg.data_ptr(), rowsums=rs.data_ptr(), nanflags=nf.data_ptr(), ksplit=1)
        assert lib.gv2_dev_gj_epilogue(ctx, n, 0, ntl, 1, C.byref(cp), C.byref(cg), None, _lib.SCHEMES[scheme], pw, out_kind, 0,
                                       vp(sep[b]), 0) == 0, lib.gv2_last_error(ctx)
    fused = torch.full((nb, n, n), -7.0, dtype=torch.float64, device="cuda")
    cgf = _lib.GjComponent(min_sums=None, rowsums=grs.data_ptr(), nanflags=gnf.data_ptr(), ksplit=1,
                           replicate_stride_rows=n, replicate_stride_nan=n)
    assert lib.gv2_dev_gj_fused(ctx, vp(gtabs), n_pad * g_pad, n, n_pad, g, g_pad, nb, C.byref(cp) if scheme == "PG" else None,
                                C.byref(cgf), _lib.SCHEMES[scheme], pw, out_kind, vp(fused), n * n) == 0, lib.gv2_last_error(ctx)
    torch.cuda.synchronize()
    a, b = fused.cpu().numpy(), sep.cpu().numpy()
    assert np.array_equal(a, np.transpose(a, (0, 2, 1)))
    # same fp32 cell arithmetic; slabs of 64 (fused) against 128 columns and one rsqrt against two divisions + sqrt
    assert np.abs(a - b).max() <= 2e-7
    zero_row = 1.0 if kind == "D" else 0.0
    assert np.all(a[:, 9, :] == zero_row) and np.all(a[:, 17, :] == zero_row if scheme == "PG" else True) and np.all(a[1, 23, :] == zero_row)
    # duplicate genomes: 1 up to the fp32 slab rounding of the tile kernels (the job's integer pair sums give exactly 1)
    assert np.all(np.abs(a[:, 3, 40] - (0.0 if kind == "D" else 1.0)) <= 2e-7)
    lib.gv2_destroy(ctx)


# ------------------------------------------------------------------------ dendrograms
@pytest.mark.parametrize("method", ["average", "complete", "weighted", "ward", "single", "centroid", "median"])
@pytest.mark.parametrize("n,kind", [(2, "random"), (3, "ties"), (17, "ties"), (130, "random"), (513, "ties"), (1500, "blocks"),
                                    (1400, "line")])
def test_device_linkage_equals_scipy(eng, n, kind, method):
    """Device linkage == scipy.cluster.hierarchy.linkage(method) row for row, bit for bit: heights from the same
    separately rounded operations, ties resolved in the order scipy's nn_chain / mst_single_linkage / fast_linkage (heap)
    resolve them.  (centroid / median on distances that are not Euclidean take square roots of negative numbers: NaN
    entries of the working matrix are never chosen, by scipy or here; rows are compared NaN == NaN.)"""
    from scipy.cluster.hierarchy import linkage
    rng = np.random.default_rng(n)
    if kind == "random":
        D = rng.random((n, n))
    elif kind == "ties":
        D = rng.integers(1, 6, (n, n)) / 5.0                       # five distinct distances: ties everywhere
    elif kind == "line":                                           # points on a line with shrinking gaps: every point's nearest
        pos = np.cumsum(1.0 + (n - np.arange(n)) * 1e-3)           # neighbour is the next one, the chain grows to n elements
        D = np.abs(pos[:, None] - pos[None, :])                    # (a worst case for the chain's bookkeeping)
    else:                                                          # block structure like a taxon table: many exact 1.0
        grp = rng.integers(0, 12, n)
        D = np.where(grp[:, None] == grp[None, :], np.round(rng.random((n, n)) * 0.6, 3), 1.0)
    D = np.triu(D, 1)
    D = D + D.T
    Z = eng.linkage(D, method)
    with np.errstate(all="ignore"):
        ref = linkage(D[np.triu_indices_from(D, k=1)], method=method)
    assert Z.shape == ref.shape and np.array_equal(Z, ref, equal_nan=True)
    # an asymmetric matrix is clustered by its upper triangle, as the reference does
    A = D + np.tril(rng.random((n, n)), -1)
    assert np.array_equal(eng.linkage(A, method), ref, equal_nan=True)


@pytest.mark.parametrize("method", ["average", "ward", "complete"])
@pytest.mark.parametrize("n,kind", [(513, "ties"), (1500, "blocks"), (2300, "random")])
def test_device_linkage_kernel_variants_agree(eng, monkeypatch, n, kind, method):
    """The three chain kernels -- eager column updates (trees too large for the shared-memory stamps), lazy columns, lazy
    columns + live list (the default) -- give the same bit-identical linkage matrix (GV2_LK_LAZY / GV2_LK_LIST are read at
    every launch)."""
    from scipy.cluster.hierarchy import linkage
    rng = np.random.default_rng(n)
    if kind == "ties":
        D = rng.integers(1, 6, (n, n)) / 5.0
    elif kind == "random":
        D = rng.random((n, n))
    else:
        grp = rng.integers(0, 12, n)
        D = np.where(grp[:, None] == grp[None, :], np.round(rng.random((n, n)) * 0.6, 3), 1.0)
    D = np.triu(D, 1)
    D = D + D.T
    ref = linkage(D[np.triu_indices_from(D, k=1)], method=method)
    for env in ({"GV2_LK_LAZY": "0"}, {"GV2_LK_LIST": "0"}, {}):
        monkeypatch.delenv("GV2_LK_LAZY", raising=False)
        monkeypatch.delenv("GV2_LK_LIST", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        assert np.array_equal(eng.linkage(D, method), ref), env


@pytest.mark.parametrize("tag", ["A", "B"])
def test_dist_mat_to_tree_golden(golden, eng, tag):
    """DistMat2Tree drop-in against the Newick string the unmodified reference printed for the same similarity matrix."""
    import gravity2_b200 as G
    S = golden[f"{tag}_S_PG_p1"]
    D = 1 - S
    D[D < 0] = 0
    names = [f"v{i}" for i in range(len(S))]
    assert G.DistMat2Tree(D.copy(), names, "average", False) == str(golden[f"{tag}_newick_avg"])
    # negative entries of the caller's matrix are zeroed in place, as dist_mat_to_tree.py:7 does
    Dn = D.copy()
    Dn[D == 0] = -0.5
    assert G.DistMat2Tree(Dn, names, "average", False) == str(golden[f"{tag}_newick_avg"]) and np.array_equal(Dn, D)
    assert G.DistMat2Tree(D.copy(), names, "complete", True) == O.dist_mat_to_tree(D.copy(), names, "complete", True)
    for method in ("centroid", "median"):          # non-monotone heights: negative branch lengths print as the reference's
        assert G.DistMat2Tree(D.copy(), names, method, False) == O.dist_mat_to_tree(D.copy(), names, method, False)
    with pytest.raises(ValueError):                # scipy's own refusal of an unknown method
        G.DistMat2Tree(D, names, "upgma", False)


def test_bootstrap_newick_equals_matrices_then_scipy(eng):
    """The device tree path (matrices clustered in HBM, only Newick lines returned) against the matrix path of the same
    engine followed by the reference's own DistMat2Tree (oracle: scipy linkage + GetNewick), replicate by replicate."""
    import gravity2_b200 as G
    t = synth.make_tables(180, 900, 6, 0.04, seed=31)
    payload = {"N_Bootstrap": 5, "SimilarityMeasurementScheme": "PG", "p": 1.0, "Dendrogram_LinkageMethod": "average",
               "PphmmNeighbourhoodWeight": 0.0, "PphmmSigScoreThreshold": 0}
    names = [f"{x}_{i}" for i, x in enumerate(t.taxo)]
    np.random.seed(77)
    mats = [m.copy() for m in G.bootstrap_distance_matrices(t.sig, None, t.taxo, t.gom_ids, payload, pipeline="pl1")]
    np.random.seed(77)
    trees = list(G.bootstrap_newick(t.sig, None, t.taxo, t.gom_ids, payload, names, pipeline="pl1"))
    assert len(trees) == 5
    for D, tree in zip(mats, trees):
        assert tree == O.dist_mat_to_tree(D, names, "average", False)
        assert tree.count(",") == len(names) - 1 and tree.endswith(");")


@pytest.mark.parametrize("n,method", [(1, "average"), (2, "average"), (5, "ward"), (70, "single"), (60, "centroid"), (33, "median")])
def test_bootstrap_newick_small_and_pl2(eng, n, method):
    """Tree path edge cases: one and two genomes, PL2 (sorted indices, real location table, GOMDB from the reference rows
    only), the other linkage methods -- always against the matrix path of the same engine + the reference's tree code."""
    import gravity2_b200 as G
    t = synth.make_tables(n, 120, min(3, n), 0.3, seed=n)
    n_ref = max(1, n - 2)
    payload = {"N_Bootstrap": 3, "SimilarityMeasurementScheme": "PG", "p": 1.0, "Dendrogram_LinkageMethod": method,
               "PphmmNeighbourhoodWeight": 0.0, "PphmmSigScoreThreshold": 0}
    names = [f"g{i}" for i in range(n)]
    np.random.seed(5)
    mats = [m.copy() for m in G.bootstrap_distance_matrices(t.sig, t.loc, t.taxo[:n_ref], t.gom_ids, payload, pipeline="pl2")]
    np.random.seed(5)
    trees = list(G.bootstrap_newick(t.sig, t.loc, t.taxo[:n_ref], t.gom_ids, payload, names, pipeline="pl2"))
    assert len(trees) == 3
    for D, tree in zip(mats, trees):
        if n == 1:
            assert tree == "g0:0.000000"            # scipy refuses an empty distance list; the single leaf is printed
        else:
            assert tree == O.dist_mat_to_tree(D, names, method, False)


def test_error_paths(eng):
    import gravity2_b200 as G
    with pytest.raises(ValueError):
        eng.similarity(np.ones((3, 4)), np.ones((2, 2)), None, "PG")
    with pytest.raises(ValueError):
        eng.similarity(np.ones((3, 4)), None, None, "XX")
    with pytest.raises(G.engine.GravityNativeError):           # multiplicity > 255 is refused, never wrapped
        w = np.zeros((8, 40), dtype=np.int32); w[:, 0] = 300
        eng.bootstrap(np.random.default_rng(0).random((5, 40)), None, None, 0, w, "P")
    assert G.GOMSignatureTable_Constructor(np.zeros((4, 6)), {}, [], {"N_CPUs": 1}).shape == (4, 0)
    S = eng.similarity(np.zeros((0, 5)), None, None, "P")
    assert S.shape == (0, 0)


def test_full_size_properties_cfg3(eng):
    """BASELINE config 3 size (10k genomes x 30k PPHMMs): size-independent properties + spot checks against the
    reference expression on sampled pairs (the oracle cannot run the full matrix in seconds)."""
    t = synth.make_tables(10000, 30000, 200, 0.01, seed=synth.SEED_BASE + 3)
    sig = t.sig
    sig[9001] = sig[17]                                   # duplicate genome
    sig[9002] = 0.0                                       # empty genome
    S = eng.similarity(sig, None, None, "P", 1.0)
    assert S.shape == (10000, 10000) and np.array_equal(S, S.T)
    d = np.diag(S)
    assert d[9002] == 0.0 and np.all(np.delete(d, 9002) == 1.0)
    assert S.min() >= 0.0 and S.max() <= 1.0 + 1e-12
    assert abs(S[17, 9001] - 1.0) < 1e-6 and np.all(S[9002] == 0.0)
    rng = np.random.default_rng(0)
    for i, j in rng.integers(0, 10000, (200, 2)):
        ref = O._gj(sig[i], sig[j])
        assert abs(S[i, j] - ref) <= 1e-6 * abs(ref) + 1e-12
    # tensor-core replicates at full size: all-ones multiplicities reproduce the base matrix; exact 1.0 for duplicates
    w = np.ones((8, 30000), dtype=np.int32)
    w[1:] = np.stack([np.bincount(rng.integers(0, 30000, 30000), minlength=30000) for _ in range(7)])
    from gravity2_b200.engine import Job
    job = Job(eng, sig, None, None, 0, "P", 1.0, distance=False)
    seen = {}

    def take(b, m):
        seen[b] = (m[:64, :].copy(), float(m[17, 9001]), float(m[9002, 9002]), bool(np.array_equal(m[:300, 5000:5300], m[5000:5300, :300].T)))
    job.for_each_replicate(w, take)
    job.close()
    assert sorted(seen) == list(range(8))
    assert np.allclose(seen[0][0], S[:64], rtol=1e-6, atol=1e-12)
    for b in range(8):
        assert seen[b][1] == 1.0 and seen[b][2] == 0.0 and seen[b][3]
    # device UPGMA at full size: the base distance matrix (10 000 leaves, thousands of exact ties at D = 1) against scipy
    from scipy.cluster.hierarchy import linkage
    D = 1.0 - S
    D[D < 0] = 0
    Zd = eng.linkage(D, "average")
    assert np.array_equal(Zd, linkage(D[np.triu_indices_from(D, k=1)], method="average"))
    i, j = 3, 4711
    ww = w[5].astype(np.float64)
    ref = np.sum(np.minimum(sig[i], sig[j]) * ww) / np.sum(np.maximum(sig[i], sig[j]) * ww)
    assert abs(seen[5][0][i, j] - ref) <= 1e-6 * ref + 1e-12
