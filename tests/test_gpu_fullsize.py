"""GPU parity at the BENCHED shapes (run with -m gpu on a B200), through the C ABI.

The bench workloads (BASELINE configs 2, 3 and 5) are checked at their own size by sampling cells of the device
results against the single-cell restatements of the reference's expressions (oracle/sampled_check.py, pinned to
the oracle's literal loop bodies by tests/test_sampled_check.py):
  * GOM signature table cells (replicate, genome, GOM) against the literally resampled dcor: rel <= 1e-9;
  * distance-matrix cells (replicate, i, j) against the reference expression on the replicate's GOM rows: abs <= 1e-6
    (north_star tolerance), incl. an own-taxon pair, a duplicate genome and an empty genome;
plus the device-RNG weights, the a16 reference golden, negative / NaN entries through the job path, the up-front
multiplicity check and the exactly degenerate dCor case.
"""
import ctypes as C
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import gravity_oracle as O
from oracle import philox_oracle
from oracle import sampled_check as SC
from gravity2_b200 import synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUP_TOL = 1e-7      # duplicate genomes: D = 0 up to the fp32 slab rounding of the GOM pair sums


@pytest.fixture(scope="module")
def eng():
    import gravity2_b200
    return gravity2_b200.default_engine()


def _weights(p, nb, seed, sort=False):
    idx = synth.reference_resample_indices(p, nb, seed=seed, sort=sort)
    return np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32)


def _check_job_sampled(eng, sig, loc, taxo_ref, gom_ids, w, n_gom_cells, pairs_per_rep, special, rng, pl1):
    """Runs the PG job (base + len(w) replicates, default replicate path) and compares sampled cells with the oracle.
    Returns (number of GOM cells checked, number of matrix cells checked, worst errors)."""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    n, p = sig.shape
    n_ref = len(taxo_ref)
    gor = G.gom_membership(taxo_ref, gom_ids)
    loc_tab = sig if pl1 else loc
    job = Job(eng, sig, loc_tab, gor, len(gom_ids), "PG", 1.0, distance=True, ring_bytes=4 * n * n * 8)
    try:
        nb = len(w)
        gom_rep = job.gom_table(w)                                  # (B, N, G): what the fused tail consumes
        gom_base = job.gom_table(None)
        members = {k: np.where(np.asarray(taxo_ref) == gid)[0] for k, gid in enumerate(gom_ids)}
        # ---- GOM cells: (replicate or base, genome, GOM); own-taxon cells are the heavy-kernel ones
        worst_g = 0.0
        cells = [(-1, 0, int(gor[0])), (0, 0, int(gor[0]))]
        while len(cells) < n_gom_cells:
            v = int(rng.integers(0, n))
            own = len(cells) % 3 == 0 and v < n_ref
            cells.append((int(rng.integers(-1, nb)), v, int(gor[v]) if own else int(rng.integers(0, len(gom_ids)))))
        for b, v, k in cells:
            ref = SC.gom_cell(loc_tab, members[k], v, None if b < 0 else w[b])
            got = gom_base[v, k] if b < 0 else gom_rep[b, v, k]
            if np.isnan(ref):                                      # sqrt of a sum that rounds below zero (dcor.py:10, quirk Q5)
                assert np.isnan(got), f"GOM cell (b={b}, v={v}, g={k}): device {got!r} oracle NaN"
                continue
            err = abs(got - ref)
            assert err <= 1e-9 * abs(ref) + 1e-12, f"GOM cell (b={b}, v={v}, g={k}): device {got!r} oracle {ref!r}"
            worst_g = max(worst_g, err / max(abs(ref), 1e-300))
        # ---- matrix cells
        stats = {"cells": 0, "worst": 0.0}
        taxon_all = np.asarray(list(taxo_ref) + ["?"] * (n - n_ref))

        def check(b, m):
            ww = None if b < 0 else w[b]
            gt = gom_base if b < 0 else gom_rep[b]
            for i, j in SC.sample_pairs(rng, n, taxon_all, pairs_per_rep, special=special):
                ref = SC.pair_cell(sig, gt[i], gt[j], i, j, ww, "PG", 1.0, True)
                err = abs(m[i, j] - ref)
                assert err <= 1e-6, f"D[{b}][{i},{j}]: device {m[i, j]!r} oracle {ref!r}"
                assert m[j, i] == m[i, j]
                stats["worst"] = max(stats["worst"], err)
                stats["cells"] += 1

        check(-1, job.base())
        seen = []
        job.for_each_replicate(w, lambda b, m: (seen.append(b), check(b, m)))
        assert seen == list(range(nb))
        return len(cells), stats["cells"], worst_g, stats["worst"]
    finally:
        job.close()


def test_cfg3_shape_pl1_pg_replicates_sampled_vs_oracle(eng):
    """BASELINE config 3 (10 000 x 30 000, G = 200, PL1 PG, sparse pair lists + weighted GOM dCor + fused tail):
    the workload bench.py times, 64 replicates on the default replicate path."""
    t = synth.make_config("cfg3", fast=True)
    sig = t.sig
    n = sig.shape[0]
    sig[7] = 0.0                                   # empty genome
    sig[n - 1] = sig[2]                            # duplicate genome (in another taxon)
    w = _weights(30000, 64, seed=77)
    rng = np.random.default_rng(3)
    own = np.where(t.taxo == t.taxo[5])[0]
    special = [(2, n - 1), (7, 11), (7, 7), (5, int(own[-1])), (n - 1, n - 1)]
    ng, nc, eg, ec = _check_job_sampled(eng, sig, None, t.taxo, t.gom_ids, w, 32, 4, special, rng, pl1=True)
    assert ng >= 32 and nc >= 200
    print(f"cfg3 shape: {ng} GOM cells (worst rel {eg:.2e}), {nc} matrix cells (worst abs {ec:.2e})")


def test_cfg3_shape_duplicate_and_empty_genomes(eng):
    """duplicate genomes: D = 0; empty genome: D == 1 against everything and on its own diagonal (S = 0)"""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    t = synth.make_tables_sparse(3000, 30000, 60, 0.01, seed=9)
    sig = t.sig
    sig[7] = 0.0
    sig[2999] = sig[2]
    gor = G.gom_membership(t.taxo, t.gom_ids)
    # genome 2999 must sit in genome 2's taxon for identical GOM rows of the two to be guaranteed by construction
    gor[2999] = gor[2]
    w = _weights(30000, 40, seed=5)
    job = Job(eng, sig, sig, gor, len(t.gom_ids), "PG", 1.0, distance=True)
    seen = {}
    job.for_each_replicate(w, lambda b, m: seen.__setitem__(b, (float(m[2, 2999]), float(m[2999, 2]), float(m[7, 7]),
                                                                  float(m[7].min()), float(m[11, 11]))))
    job.close()
    for b in range(40):
        d_dup, d_dup_t, d_empty_diag, d_empty_min, d_diag = seen[b]
        assert d_dup <= DUP_TOL and d_dup_t == d_dup and d_diag == 0.0
        assert d_empty_diag == 1.0 and d_empty_min == 1.0


def test_cfg5_shape_pl2_sampled_vs_oracle(eng):
    """BASELINE config 5 (pipeline_ii update): 20 000 reference + 5 000 query genomes x 30 000 PPHMMs, 200 GOMs built
    from the reference rows only, sorted resample indices, real location table (virus_classification.py:290-315)."""
    t = synth.make_config("cfg5", fast=True)
    n_ref = 20000
    w = _weights(30000, 4, seed=78, sort=True)
    rng = np.random.default_rng(5)
    special = [(0, 24999), (20001, 3), (24000, 24000)]
    ng, nc, eg, ec = _check_job_sampled(eng, t.sig, t.loc, t.taxo[:n_ref], t.gom_ids, w, 24, 12, special, rng, pl1=False)
    assert ng >= 24 and nc >= 60
    print(f"cfg5 shape: {ng} GOM cells (worst rel {eg:.2e}), {nc} matrix cells (worst abs {ec:.2e})")


def test_cfg2_with_replicates_full_matrices(eng):
    """BASELINE config 2 (1 000 x 5 000, G = 50, 100 replicates): every replicate streamed; three whole replicate
    matrices against the oracle expression on the (sample-verified) device GOM tables, GOM cells sampled."""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    t = synth.make_config("cfg2")
    n, p = t.sig.shape
    w = _weights(p, 100, seed=79)
    gor = G.gom_membership(t.taxo, t.gom_ids)
    job = Job(eng, t.sig, t.sig, gor, len(t.gom_ids), "PG", 1.0, distance=True)
    gom_rep = job.gom_table(w)
    keep = {}
    job.for_each_replicate(w, lambda b, m: keep.__setitem__(b, m.copy()) if b in (0, 41, 99) else None)
    job.close()
    rng = np.random.default_rng(2)
    for _ in range(24):
        b, v, k = int(rng.integers(0, 100)), int(rng.integers(0, n)), int(rng.integers(0, len(t.gom_ids)))
        ref = SC.gom_cell(t.sig, np.where(t.taxo == t.gom_ids[k])[0], v, w[b])
        assert abs(gom_rep[b, v, k] - ref) <= 1e-9 * abs(ref) + 1e-12
    for b in (0, 41, 99):
        ref = O.dist_from_sim(O.similarity_mat_fast(t.sig, gom_rep[b], None, 0, "PG", 1.0, w=w[b]))
        assert np.abs(keep[b] - ref).max() <= 1e-6
        assert np.array_equal(keep[b], keep[b].T)


def test_dev_philox_weights_bit_exact(eng):
    """gv2_dev_philox_weights: multiplicities of the counter-based resample == bincount of the Random123 restatement,
    for a replicate offset b0 > 0 as a multi-GPU rank would ask for"""
    lib, ctx = eng.lib, eng.ctx
    from gravity2_b200 import _lib
    for seed, b0, nb, p in [(0, 0, 2, 7), (0xDEADBEEFCAFEF00D, 3, 5, 1001), (42, 996, 4, 30000)]:
        d = C.c_void_p()
        _lib.check(ctx, lib.gv2_dev_alloc(ctx, nb * p * 4, C.byref(d)))
        try:
            _lib.check(ctx, lib.gv2_dev_philox_weights(ctx, C.c_uint64(seed), b0, nb, p, d))
            got = np.empty((nb, p), dtype=np.int32)
            _lib.check(ctx, lib.gv2_memcpy_d2h(ctx, got.ctypes.data_as(C.c_void_p), d, got.nbytes))
        finally:
            lib.gv2_dev_free(ctx, d)
        idx = philox_oracle.philox_indices(seed, b0 + nb, p)[b0:]
        ref = np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32)
        assert np.array_equal(got, ref)
        assert np.all(got.sum(axis=1) == p)


@pytest.mark.parametrize("tag", ["E", "F"])
def test_sort_pphmm_reference_golden(eng, tag):
    """a16 against the UNMODIFIED RefVirusAnnotator.sort_pphmm (tests/golden/reference_golden_sort_pphmm.npz)"""
    import gravity2_b200 as G
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "reference_golden_sort_pphmm.npz"), allow_pickle=False))
    D = G.sort_pphmm_distances(g[f"{tag}_sig"])
    ref = g[f"{tag}_traits_D"]
    assert np.array_equal(np.isnan(D), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(D, nan=-1.0), np.nan_to_num(ref, nan=-1.0))      # popcount ratio: bit for bit
    if tag == "E":
        assert G.sort_pphmm_tree(g["E_sig"]) == str(g["E_traits_newick"])
    else:
        with pytest.raises(ValueError):
            G.sort_pphmm_tree(g["F_sig"])


def test_job_negative_and_nan_entries(eng, monkeypatch):
    """threshold 0 zeroes negative scores on every call of the reference (similarity_matrix_constructor.py:139-140) and
    a NaN score zeroes its genome's row only in the replicates that drew its column -- through the device job, on all
    three replicate paths"""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    t = synth.make_tables(90, 600, 4, 0.1, seed=12)
    sig = t.sig
    rng = np.random.default_rng(6)
    hit = np.argwhere(sig != 0)
    for r, c in hit[rng.choice(len(hit), 60, replace=False)]:
        sig[r, c] = -sig[r, c]                                     # negative scores
    sig[13, 100] = np.nan
    sig[40, 7] = np.nan
    w = np.stack([np.bincount(rng.integers(0, 600, 600), minlength=600) for _ in range(12)]).astype(np.int32)
    w[0, 100] = 0; w[1, 100] = 2; w[2, 7] = 0; w[3, 7] = 1
    ref_base = O.similarity_mat(sig.copy(), None, None, None, 0.0, 0, "P", 1.0)
    for path in ("sparse", "mma", "fp32"):
        monkeypatch.setenv("GV2_BOOT_PATH", path)
        job = Job(eng, sig, None, None, 0, "P", 1.0, distance=False)
        base = job.base()
        reps = {}
        job.for_each_replicate(w, lambda b, m: reps.__setitem__(b, m.copy()))
        job.close()
        assert np.abs(base - ref_base).max() <= 1e-6, path
        for b in range(12):
            ref = O.similarity_mat(sig[:, np.repeat(np.arange(600), w[b])].copy(), None, None, None, 0.0, 0, "P", 1.0)
            assert np.abs(reps[b] - ref).max() <= 1e-6, (path, b)
        assert np.all(reps[1][13] == 0.0) and reps[0][13, 13] == 1.0      # NaN column drawn / not drawn
        assert np.all(reps[3][40] == 0.0) and reps[2][40, 40] == 1.0
    monkeypatch.delenv("GV2_BOOT_PATH")
    # the PG job keeps the location side un-thresholded (PL1: the signature table doubles as location table)
    sig2 = np.nan_to_num(sig, nan=0.0)
    gor = G.gom_membership(t.taxo, t.gom_ids)
    job = Job(eng, sig2, sig2, gor, len(t.gom_ids), "PG", 1.0, distance=True)
    D0 = job.base()
    job.close()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gom = O.gom_table(sig2, O.gomdb(t.taxo, sig2, t.gom_ids), t.gom_ids)
        ref = O.dist_from_sim(O.similarity_mat(sig2.copy(), gom, sig2, None, 0.0, 0, "PG", 1.0))
    assert np.abs(D0 - ref).max() <= 1e-6


def test_weight_overflow_refused_before_delivery_and_flag_cleared(eng):
    """a multiplicity above 255 (uint8 operands) is refused BEFORE anything is delivered, and the next call on the same
    job works (the flag does not stick)"""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    t = synth.make_tables(70, 500, 4, 0.1, seed=13)
    gor = G.gom_membership(t.taxo, t.gom_ids)
    rng = np.random.default_rng(1)
    good = np.stack([np.bincount(rng.integers(0, 500, 500), minlength=500) for _ in range(40)]).astype(np.int32)
    bad = good.copy()
    bad[37, 3] = 300
    job = Job(eng, t.sig, t.sig, gor, len(t.gom_ids), "PG", 1.0, distance=True, ring_bytes=4 * 70 * 70 * 8)
    try:
        delivered = []
        with pytest.raises(G.engine.GravityNativeError):
            job.for_each_replicate(bad, lambda b, m: delivered.append(b))
        assert delivered == []
        with pytest.raises(G.engine.GravityNativeError):
            job.gom_table(bad)
        out = {}
        job.for_each_replicate(good, lambda b, m: out.__setitem__(b, m.copy()))
        assert sorted(out) == list(range(40))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = O.pl1_bootstrap_replicate(t.sig, t.taxo, t.gom_ids, np.repeat(np.arange(500), good[39]))
        assert np.abs(out[39] - ref).max() <= 1e-6
    finally:
        job.close()
    # host delivery of several replicates through a one-slot ring is refused (the two halves ping-pong)
    lib, ctx = eng.lib, eng.ctx
    from gravity2_b200 import _lib
    job = Job(eng, t.sig, None, None, 0, "P", 1.0, distance=True)
    d_w, d_ring = C.c_void_p(), C.c_void_p()
    host = np.empty((2, 70, 70))
    _lib.check(ctx, lib.gv2_dev_alloc(ctx, good.nbytes, C.byref(d_w)))
    _lib.check(ctx, lib.gv2_dev_alloc(ctx, 70 * 70 * 8, C.byref(d_ring)))
    _lib.check(ctx, lib.gv2_memcpy_h2d(ctx, d_w, good.ctypes.data_as(C.c_void_p), good.nbytes))
    rc = lib.gv2_job_replicates_stream(job.job, d_w, 2, d_ring, 1, host.ctypes.data_as(C.c_void_p), None, None)
    assert rc == _lib.ERR_ARG
    lib.gv2_dev_free(ctx, d_w)
    lib.gv2_dev_free(ctx, d_ring)
    job.close()


def test_degenerate_dcor_constant_location_values(eng):
    """A genome whose drawn hits all carry ONE location value and cover the GOM's whole column set (a single-member
    taxon in PL1, where the score table doubles as location table): the reference's centred matrix is exactly zero and
    dcor returns 0.0 (dcor.py:54-60) -- not a ratio of rounding noise."""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    t = synth.make_tables(50, 400, 5, 0.1, seed=14)
    sig = t.sig
    taxo = t.taxo.copy()
    taxo[49] = "SOLO"
    sig[49] = 0.0
    sig[49, [3, 50, 51, 200, 333, 399]] = 37.3            # six hits, one value
    sig[48] = 0.0
    sig[48, [10, 20]] = [5.5, 5.5]                        # member of a normal taxon with constant values
    ids = list(t.gom_ids) + ["SOLO"]
    gor = G.gom_membership(taxo, ids)
    rng = np.random.default_rng(3)
    w = np.stack([np.bincount(rng.integers(0, 400, 400), minlength=400) for _ in range(33)]).astype(np.int32)
    job = Job(eng, sig, sig, gor, len(ids), "PG", 1.0, distance=True)
    try:
        g0 = job.gom_table(None)
        gw = job.gom_table(w)
    finally:
        job.close()
    solo = len(ids) - 1
    assert g0[49, solo] == 0.0 and np.all(gw[:, 49, solo] == 0.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref0 = O.gom_table(sig, O.gomdb(taxo, sig, ids), ids)
        assert np.allclose(g0, ref0, rtol=1e-9, atol=1e-12, equal_nan=True)
        for b in (0, 32):
            idx = np.repeat(np.arange(400), w[b])
            bl = sig[:, idx]
            ref = O.gom_table(bl, O.gomdb(taxo, bl, ids), ids)
            assert np.allclose(gw[b], ref, rtol=1e-9, atol=1e-12, equal_nan=True)


# ------------------------------------------------------------------ next-2 / next-3 rows of SURVEY 8f
def test_classification_epilogue_on_device_matrix(eng):
    """PairwiseSimilarityScore_Cutoff_Dict_Constructor + the 1-NN proposal on a matrix that stays in HBM == the same functions
    on the host copy == the reference golden; and the batched PL2 helper against per-replicate host evaluation."""
    import gravity2_b200 as G
    from gravity2_b200 import classification as CL
    from gravity2_b200.engine import Job
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "reference_golden_classification.npz"), allow_pickle=False))
    S, taxo, n_pair = g["H_S"], g["H_taxo"], int(g["H_npair"])
    import torch
    d_S = torch.from_numpy(S).cuda()
    dm = CL.DeviceMatrix(eng, d_S.data_ptr(), S.shape[0])
    np.random.seed(int(g["H_seed"]))
    d = CL.PairwiseSimilarityScore_Cutoff_Dict_Constructor(dm, taxo, n_pair)
    for cls, v in d.items():
        assert np.array_equal(np.array(v["PairwiseSimilarityScore_IntraClass_List"]), g[f"H_{cls}_intra"])
        assert np.array_equal(np.array(v["PairwiseSimilarityScore_InterClass_List"]), g[f"H_{cls}_inter"])
        assert v["CutOff"] == float(g[f"H_{cls}_cutoff"])
    assert np.random.random() == float(g["H_rng_next"])
    mx, arg = dm.rowmax(40, 60, 0, 40)
    assert np.array_equal(mx, S[40:, :40].max(axis=1)) and np.array_equal(arg, S[40:, :40].argmax(axis=1))
    S2 = S.copy(); S2[45, 3] = S2[45, 17] = 5.0; S2[50, 9] = np.nan
    d_S2 = torch.from_numpy(S2).cuda()
    mx, arg = CL.DeviceMatrix(eng, d_S2.data_ptr(), 60).rowmax(40, 60, 0, 40)
    assert arg[5] == 3 and arg[10] == 9 and np.isnan(mx[10])             # first maximum; NaN wins, as numpy
    # batched PL2 helper: device-resident similarity, cells and row maxima only
    t = synth.make_tables(90, 500, 5, 0.1, seed=15)
    n_ref, n_ucf = 70, 20
    payload = {"N_Bootstrap": 3, "SimilarityMeasurementScheme": "PG", "p": 1.0, "N_PairwiseSimilarityScores": 150}
    np.random.seed(8)
    got = list(G.bootstrap_classification(t.sig, t.loc, t.taxo[:n_ref], t.gom_ids, payload, n_ucf))
    np.random.seed(8)
    gor = G.gom_membership(t.taxo[:n_ref], t.gom_ids)
    for b in range(3):
        idx = G.draw_resample_indices(500, sort=True)
        w = G.weights_from_indices(idx, 500)
        Sb = eng.bootstrap(t.sig, t.loc, gor, len(t.gom_ids), w[None, :], "PG", 1.0, distance=False)[0]
        ref_d = CL.PairwiseSimilarityScore_Cutoff_Dict_Constructor(Sb[:n_ref][:, :n_ref], t.taxo[:n_ref], 150)
        ref_mx, ref_tx = CL.max_similarity(Sb[-n_ucf:][:, :n_ref], t.taxo[:n_ref])
        d_b, mx_b, tx_b = got[b]
        assert list(d_b) == list(ref_d)
        for cls in ref_d:
            assert d_b[cls]["PairwiseSimilarityScore_InterClass_List"] == ref_d[cls]["PairwiseSimilarityScore_InterClass_List"]
            assert d_b[cls]["PairwiseSimilarityScore_IntraClass_List"] == ref_d[cls]["PairwiseSimilarityScore_IntraClass_List"]
            assert d_b[cls]["CutOff"] == ref_d[cls]["CutOff"]
        assert np.array_equal(mx_b, ref_mx) and np.array_equal(tx_b, ref_tx)


@pytest.mark.parametrize("tag", ["J", "K"])
def test_virus_grouping_estimator_reference_golden(eng, tag):
    import gravity2_b200 as G
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "reference_golden_grouping.npz"), allow_pickle=False))
    groups, cut, score, u1, u2 = G.VirusGrouping_Estimator(g[f"{tag}_D"], "average", g[f"{tag}_taxo"])
    assert np.array_equal(groups, g[f"{tag}_groups"])
    assert np.array_equal(np.array([cut, score, u1, u2]), g[f"{tag}_stats"])


def test_braycurtis_matrix_vs_sklearn(eng, tmp_path):
    """pphmm_loc_diffs_pairwise (app/utils/shared_pphmm_graphs.py:202-207): the reference's own call,
    sklearn pairwise_distances(..., metric="braycurtis"), on the location CSV laid out as pl1_graphs.py:317-319 writes it"""
    import pandas as pd
    from sklearn.metrics.pairwise import pairwise_distances
    import gravity2_b200 as G
    t = synth.make_tables(70, 900, 5, 0.08, seed=16)
    naive = np.abs(t.loc)
    naive[11] = 0.0; naive[29] = 0.0                                    # two genomes without any hit: 0 / 0
    cols = [f"PPHMM {c}" for c in range(900)]
    df = pd.DataFrame(naive, columns=cols)
    df.insert(0, "Virus name", [f"v{i}" for i in range(70)])
    f = os.path.join(tmp_path, "locs.csv")
    df.to_csv(f, index=False)
    order = np.random.default_rng(0).permutation(70)
    got = G.pphmm_loc_diffs_pairwise({"PphmmLocs": f}, order, None)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = pairwise_distances(np.nan_to_num(naive[order], nan=0), metric="braycurtis")
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] - ref[ok]).max() <= 1e-6


def test_native_errors_follow_the_reference_convention(eng):
    """the modules mirroring the reference's import paths end a run the way the reference does:
    raise_gravity_error -> SystemExit("FATAL ERROR: ...") (app/utils/error_handlers.py:3-4)"""
    from gravity2_b200.app.utils.parallel_gom_sig_generator import GOMSignatureTable_Constructor
    from gravity2_b200.app.utils.shared_pphmm_graphs import shared_norm_pphmm_ratio
    import gravity2_b200 as G
    loc = np.zeros((4, 6)); loc[:, 0] = [1.0, 2.0, 3.0, 4.0]
    db = {"A": np.full((2, 7), 1.0)}                                   # GOM rows of the wrong width: native argument error
    with pytest.raises((SystemExit, ValueError)) as ei:
        GOMSignatureTable_Constructor(loc, db, ["A"], {"N_CPUs": 1})
    if ei.type is SystemExit:
        assert "FATAL ERROR" in str(ei.value)
    with pytest.raises(ZeroDivisionError):                              # the reference's own Python error passes through
        G.shared_norm_pphmm_ratio_from_array(np.zeros((3, 5)))


# ------------------------------------------------------------------ every GPU of the node behind one caller
def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_multi_engine_single_process_equals_one_gpu(eng):
    """gv2_multi_*: one process, one host thread per GPU inside the library.  With one visible GPU the same entry points run
    on a one-device set; with several, replicates are dealt by index and the base matrix is tile-sharded + all-gathered (NCCL
    when libnccl.so.2 can be dlopen'ed).  Results must be bit-identical to the single-GPU job."""
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    ndev = min(_n_gpus(), 8)
    meng = G.MultiEngine(list(range(ndev)))
    assert meng.n_devices == ndev
    t = synth.make_tables(300, 1200, 6, 0.05, seed=17)
    gor = G.gom_membership(t.taxo, t.gom_ids)
    w = _weights(1200, 21, seed=3)
    names = [f"v{i}" for i in range(300)]
    one = Job(eng, t.sig, t.sig, gor, len(t.gom_ids), "PG", 1.0, distance=True)
    base1 = one.base()
    trees1, mats1 = {}, {}
    one.for_each_tree(w, names, lambda b, s: trees1.__setitem__(b, s))
    one.for_each_replicate(w, lambda b, m: mats1.__setitem__(b, m.copy()))
    one.close()
    mj = G.MultiJob(meng, t.sig, t.sig, gor, len(t.gom_ids), "PG", 1.0, distance=True)
    try:
        assert np.array_equal(mj.base(), base1)
        order, trees, mats = [], {}, {}
        mj.for_each_tree(w, names, lambda b, s: (order.append(b), trees.__setitem__(b, s)))
        assert order == list(range(21)) and trees == trees1
        order2 = []
        mj.for_each_replicate(w, lambda b, m: (order2.append(b), mats.__setitem__(b, m.copy())))
        assert order2 == list(range(21))
        for b in range(21):
            assert np.array_equal(mats[b], mats1[b])
        gen = [m.copy() for m in mj.replicates(w[:5])]
        assert all(np.array_equal(gen[b], mats1[b]) for b in range(5))
    finally:
        mj.close()
    # the batched drop-in on every GPU: same lines as on one
    payload = {"N_Bootstrap": 7, "SimilarityMeasurementScheme": "PG", "p": 1.0}
    np.random.seed(5)
    a = list(G.bootstrap_newick(t.sig, None, t.taxo, t.gom_ids, payload, names, pipeline="pl1", engine=eng))
    np.random.seed(5)
    b = list(G.bootstrap_newick(t.sig, None, t.taxo, t.gom_ids, payload, names, pipeline="pl1", engine=meng))
    assert a == b
    print(f"multi engine: {ndev} device(s), nccl={meng.uses_nccl}")
    meng.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two GPUs")
def test_multi_engine_uses_nccl_allgather_on_two_or_more_gpus(eng):
    import gravity2_b200 as G
    meng = G.MultiEngine(None)
    assert meng.n_devices >= 2 and meng.uses_nccl
    t = synth.make_tables(700, 2000, 8, 0.03, seed=18)
    gor = G.gom_membership(t.taxo, t.gom_ids)
    mj = G.MultiJob(meng, t.sig, t.loc, gor, len(t.gom_ids), "PG", 1.0, distance=False)
    try:
        S = mj.base()
    finally:
        mj.close()
    ref = eng.similarity(t.sig, eng.gom_table(t.loc, t.loc[np.argsort(gor, kind="stable")],
                                              np.concatenate([[0], np.cumsum(np.bincount(gor, minlength=len(t.gom_ids)))])), None, "PG", 1.0)
    assert np.abs(S - ref).max() <= 1e-6 and np.array_equal(S, S.T)
    meng.close()


def test_sparse_coo_tables_equal_dense_tables(eng, tmp_path):
    """the *_coo tables the reference pickles (app/src/ref_virus_annotator.py:395-400) go to the device as
    triplets; everything downstream is bit-identical to the dense upload"""
    import pickle
    from scipy.sparse import coo_matrix
    import gravity2_b200 as G
    from gravity2_b200.engine import Job
    t = synth.make_tables(200, 1500, 6, 0.04, seed=19)
    gor = G.gom_membership(t.taxo, t.gom_ids)
    w = _weights(1500, 9, seed=4)
    f = os.path.join(tmp_path, "RefVirusAnnotator.p")
    with open(f, "wb") as fh:       # the keys pl1 writes
        pickle.dump({"PPHMMSignatureTable": t.sig, "PPHMMLocationTable": t.loc, "GOMIDList": t.gom_ids,
                     "PPHMMSignatureTable_coo": coo_matrix(t.sig), "PPHMMLocationTable_coo": coo_matrix(t.loc)}, fh)
    tabs = G.load_annotator_tables(f)
    assert tabs["PPHMMSignatureTable"].nnz == np.count_nonzero(t.sig)
    out = {}
    for kind, (sg, lc) in {"dense": (t.sig, t.loc), "coo": (tabs["PPHMMSignatureTable"], tabs["PPHMMLocationTable"])}.items():
        job = Job(eng, sg, lc, gor, len(t.gom_ids), "PG", 1.0, distance=True)
        reps = {}
        job.for_each_replicate(w, lambda b, m: reps.__setitem__(b, m.copy()))
        out[kind] = (job.base(), reps)
        job.close()
    assert np.array_equal(out["dense"][0], out["coo"][0])
    for b in range(9):
        assert np.array_equal(out["dense"][1][b], out["coo"][1][b])
    payload = {"N_Bootstrap": 3, "SimilarityMeasurementScheme": "PG", "p": 1.0}
    names = [f"v{i}" for i in range(200)]
    np.random.seed(2)
    a = list(G.bootstrap_newick(t.sig, None, t.taxo, t.gom_ids, payload, names, pipeline="pl1"))
    np.random.seed(2)
    b = list(G.bootstrap_newick(tabs["PPHMMSignatureTable"], None, t.taxo, t.gom_ids, payload, names, pipeline="pl1"))
    assert a == b
