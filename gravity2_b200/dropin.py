"""Drop-in replacements for the reference's hot-path functions (same names, argument order,
argument meaning, return shapes/dtypes and error behaviour), computing on a B200 through
libgravity_b200.  Import paths mirroring the reference's live under ``gravity2_b200.app.utils``.

Reference (Mayne941/gravity2) interfaces replaced here:
  SimilarityMat_Constructor        app/utils/similarity_matrix_constructor.py:7-226
  GOMDB_Constructor                app/utils/gomdb_constructor.py:2-8
  GOMSignatureTable_Constructor    app/utils/parallel_gom_sig_generator.py:10-28 (5-arg) and
                                   app/utils/gom_signature_table_constructor.py:34-52 (4-arg)
  generate_gom_sigs                app/utils/parallel_gom_sig_generator.py:30-37
  dcor                             app/utils/dcor.py:34-62
  shared_pphmm_ratio / shared_norm_pphmm_ratio   app/utils/shared_pphmm_graphs.py:59-97
  bootstrap loop bodies            app/src/pl1_graphs.py:78-107, app/src/virus_classification.py:290-315
"""
from __future__ import annotations

from typing import Iterator, Optional, Sequence

import numpy as np

from .engine import Engine, default_engine


# ------------------------------------------------------------------------------ similarity
def SimilarityMat_Constructor(PPHMMSignatureTable, GOMSignatureTable, PPHMMLocationTable, SPRSignatureTable,
                              pphmm_neighbourhood_weight, pphmm_signature_score_threshold,
                              SimilarityMeasurementScheme="PG", p=1.0, fnames=False, engine: Optional[Engine] = None):
    """Similarity matrix for the given scheme and p: float64 (N, N), fresh array.

    ``pphmm_neighbourhood_weight == 0`` (the default and every shipped payload): ``fnames`` is not read --
    the reference re-reads two CSVs from it to build neighbourhood index lists that then have no effect
    (SURVEY quirk Q2).  With a non-zero weight the lists are built from ``fnames`` exactly as the
    reference builds them and its visit-count dependent in-place row scaling (quirk Q1) is reproduced: at
    pair (i, j >= i) row i has been scaled j + 2 times and row j i + 1 times, and the caller's table is
    left scaled N + 1 times, so that a following call (e.g. the next bootstrap) starts where the
    reference's would."""
    scheme = SimilarityMeasurementScheme
    eng = engine or default_engine()
    thr = pphmm_signature_score_threshold
    aux = None
    if "R" in scheme:
        aux = SPRSignatureTable
    if "L" in scheme:
        aux = eng.location_dcor_matrix(PPHMMLocationTable)
    if pphmm_neighbourhood_weight != 0 and "P" in scheme:
        from . import neighbourhood as nbhd
        if not fnames:
            raise ValueError("PphmmNeighbourhoodWeight != 0 needs fnames['PphmmLocs'] and fnames['PphmmAndGomSigs']")
        cls = nbhd.neighbourhood_classes(*nbhd.read_neighbourhood_csvs(fnames))
        n = PPHMMSignatureTable.shape[0]
        if cls.shape != PPHMMSignatureTable.shape:
            raise IndexError(f"neighbourhood indices from the CSVs ({cls.shape}) do not fit the signature table "
                             f"{PPHMMSignatureTable.shape}")
        # cleaned GJ_P (times the R / L factor of PR / PL), then the ordinary combine with GJ_P in the aux slot:
        # P -> R, PG -> RG; PR = (R * P) ** 0.5 and PL = P * L go through R with the product as aux
        if scheme in ("PR", "PL"):
            # L is cleaned like P (:201-205) before the product; R is taken as is
            mul = np.asarray(aux, dtype=np.float64)
            if scheme == "PL":
                mul = np.where(np.isnan(mul) | (mul < 0), 0.0, mul)
            gjp = eng.similarity_neighbourhood(PPHMMSignatureTable, cls, pphmm_neighbourhood_weight, thr, mul)
            S = eng.similarity(None, None, gjp, "R", float(p) * (0.5 if scheme == "PR" else 1.0))
        else:
            gjp = eng.similarity_neighbourhood(PPHMMSignatureTable, cls, pphmm_neighbourhood_weight, thr)
            S = eng.similarity(None, GOMSignatureTable, gjp, {"P": "R", "PG": "RG"}[scheme], float(p))
        if isinstance(PPHMMSignatureTable, np.ndarray):
            nbhd.apply_in_place(PPHMMSignatureTable, cls, pphmm_neighbourhood_weight, thr, n + 1)
        return S
    S = eng.similarity(PPHMMSignatureTable, GOMSignatureTable, aux, scheme, float(p), float(thr),
                       apply_threshold=("P" in scheme), distance=False)
    if "P" in scheme and isinstance(PPHMMSignatureTable, np.ndarray):
        # the reference zeroes `row < threshold` in the CALLER's table (views, :116-140); keep that side effect
        PPHMMSignatureTable[PPHMMSignatureTable < thr] = 0
    return S


def dist_mat_constructor(PPHMMSignatureTable, GOMSignatureTable, PPHMMLocationTable, SPRSignatureTable, payload,
                         fnames=None, engine: Optional[Engine] = None):
    """GRAViTyDendrogramAndHeatmapConstruction.dist_mat_constructor, app/src/pl1_graphs.py:40-54."""
    SimMat = SimilarityMat_Constructor(PPHMMSignatureTable, GOMSignatureTable, PPHMMLocationTable, SPRSignatureTable,
                                       payload["PphmmNeighbourhoodWeight"], payload["PphmmSigScoreThreshold"],
                                       payload["SimilarityMeasurementScheme"], payload["p"], fnames, engine=engine)
    DistMat = 1 - SimMat
    DistMat[DistMat < 0] = 0
    return DistMat


# ------------------------------------------------------------------------------------ GOM
def GOMDB_Constructor(TaxoGroupingList, PPHMMLocationTable, GOMIDList):
    """GOM database: dict id -> member rows.  A host-side gather (numpy boolean mask)."""
    GOMDb = {}
    for id in GOMIDList:
        GOMDb[id] = PPHMMLocationTable[TaxoGroupingList == id, :]
    return GOMDb


def _gom_rows(GOMDB, ids, p):
    rows = [np.asarray(GOMDB[g], dtype=np.float64).reshape(-1, p) for g in ids]
    offs = np.zeros(len(ids) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([r.shape[0] for r in rows])
    cat = np.concatenate(rows, axis=0) if rows and offs[-1] > 0 else np.zeros((0, p))
    return cat, offs


def GOMSignatureTable_Constructor(PPHMMLocationTable, GOMDB, GOMIDList, payload=None, bootstrap=0,
                                  engine: Optional[Engine] = None):
    """(N, G) float64 GOM signature table, columns in GOMIDList order.  Accepts both reference
    signatures (4-arg twin without ``payload``; ``payload['N_CPUs']`` is ignored)."""
    if not isinstance(payload, (dict, type(None))):       # 4-arg twin: (loc, GOMDB, ids, bootstrap)
        bootstrap, payload = payload, None
    if bootstrap != 0:
        print(f"- (Re-)Constructing GOM Signature Table, bootstrap iteration: {bootstrap+1}")
    loc = np.asarray(PPHMMLocationTable, dtype=np.float64)
    if len(GOMIDList) == 0:
        return np.empty((len(loc), 0))
    cat, offs = _gom_rows(GOMDB, list(GOMIDList), loc.shape[1])
    return (engine or default_engine()).gom_table(loc, cat, offs)


def generate_gom_sigs(GOM, PPHMMLocationTable, GOMDB, engine: Optional[Engine] = None):
    """One GOM column as a list of floats."""
    loc = np.asarray(PPHMMLocationTable, dtype=np.float64)
    cat, offs = _gom_rows(GOMDB, [GOM], loc.shape[1])
    return [float(x) for x in (engine or default_engine()).gom_table(loc, cat, offs)[:, 0]]


def dcor(X, Y, engine: Optional[Engine] = None):
    """Distance correlation of X (n, m) and Y (n, 1) -- KAT dcor([1..5],[1,2,9,4,4]) = 0.76267624241686649."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    assert X.shape[0] == Y.shape[0]
    if Y.ndim != 2 or Y.shape[1] != 1:
        raise NotImplementedError("device dcor takes Y of shape (n, 1), as every reference call site passes")
    n = X.shape[0]
    if n == 0:
        return 0.0
    rows = np.ascontiguousarray(X.reshape(n, -1).T)
    out = (engine or default_engine()).gom_table(Y.reshape(1, n), rows, np.array([0, rows.shape[0]]), all_columns=True)
    return float(out[0, 0])


# ------------------------------------------------------------------------- presence / absence
def _sig_array(fnames, label_order):
    """get_sig_data, app/utils/shared_pphmm_graphs.py:16-21."""
    import pandas as pd
    df = pd.read_csv(fnames["PphmmAndGomSigs"], index_col=0)
    df = df.iloc[:, 1:]
    df = df.apply(pd.to_numeric)
    return np.asarray(df.reindex(label_order))


def shared_pphmm_counts(arr, engine: Optional[Engine] = None):
    return (engine or default_engine()).shared_counts(arr)


def shared_pphmm_ratio(label_order, fnames, labels, engine: Optional[Engine] = None):
    """Matrix of PPHMM co-occurrence counts: list of lists of Python ints."""
    inter, _ = shared_pphmm_counts(_sig_array(fnames, label_order), engine)
    return inter.tolist()


def shared_norm_pphmm_ratio_from_array(arr, engine: Optional[Engine] = None):
    inter, nnz = shared_pphmm_counts(arr, engine)
    mp = np.minimum(nnz[:, None], nnz[None, :]).astype(np.int64)
    hits = inter.astype(np.int64)
    lo, hi = np.minimum(hits, mp), np.maximum(hits, mp)
    if (hi == 0).any():
        raise ZeroDivisionError("division by zero")        # Python ints in the reference, :86
    return (lo / hi).T.copy()


def shared_norm_pphmm_ratio(label_order, fnames, labels, engine: Optional[Engine] = None):
    """Normalised ratio of shared PPHMMs, float64 (N, N); 0/0 raises ZeroDivisionError as the reference."""
    return shared_norm_pphmm_ratio_from_array(_sig_array(fnames, label_order), engine)


def binary_jaccard(trait_table, engine: Optional[Engine] = None):
    """PPHMM x PPHMM binary Jaccard of RefVirusAnnotator.sort_pphmm (app/src/ref_virus_annotator.py:148-158):
    popcount(a & b) / popcount(a | b), NaN for two empty rows (no zero guard in the reference)."""
    inter, nnz = shared_pphmm_counts(trait_table, engine)
    union = nnz[:, None].astype(np.int64) + nnz[None, :] - inter
    with np.errstate(invalid="ignore", divide="ignore"):
        return inter / union


def braycurtis_matrix(arr, engine: Optional[Engine] = None):
    """sklearn ``pairwise_distances(arr, metric="braycurtis")`` for a NON-NEGATIVE table: sum|u - v| / sum|u + v|.
    For u, v >= 0 that is (1 - J) / (1 + J) with J = sum(min) / sum(max) the generalised Jaccard of the similarity engine
    (sum|u - v| = sum(max) - sum(min), sum(u + v) = sum(max) + sum(min)), so the all-pairs work is the (min, +) tile
    kernel; two all-zero rows give 0 / 0 = NaN as scipy does."""
    A = np.ascontiguousarray(arr, dtype=np.float64)
    if A.size and np.nanmin(A) < 0:
        raise NotImplementedError("braycurtis_matrix is built for non-negative tables (PPHMM location distances)")
    J = (engine or default_engine()).similarity(A, None, None, "P", 1.0)
    out = (1.0 - J) / (1.0 + J)
    empty = ~(A != 0).any(axis=1)
    if empty.any():
        out[np.ix_(empty, empty)] = np.nan
    np.fill_diagonal(out, 0.0)                      # pairwise_distances pins the diagonal to 0, also for all-zero rows
    return out


def pphmm_loc_diffs_pairwise(fnames, label_order, labels=None, engine: Optional[Engine] = None):
    """app/utils/shared_pphmm_graphs.py:202-207: Bray-Curtis matrix of the genomes' PPHMM location rows.  The reference
    takes ``distance_arr`` from get_dist_data (:138-162): the location CSV with zeros -> NaN, columns permuted by median
    position (irrelevant for a pairwise sum), rows in ``label_order``; then ``nan_to_num`` -> 0."""
    import pandas as pd
    loc_df = pd.read_csv(fnames["PphmmLocs"], index_col=False)
    arr = np.asarray(loc_df.iloc[:, 1:].apply(pd.to_numeric), dtype=np.float64)
    arr = np.nan_to_num(arr, nan=0.0)[np.asarray(label_order)]
    return braycurtis_matrix(arr, engine)


# ---------------------------------------------------------------------- virus grouping
def _conditional_entropy(x, y):
    """virus_grouping_estimator.py:16-36 (summation in first-occurrence order, as collections.Counter iterates)."""
    from collections import Counter
    y_counter = Counter(y)
    xy_counter = Counter(list(zip(x, y)))
    total = float(sum(y_counter.values()))
    h = 0.0
    for xy in xy_counter.keys():
        p_xy = xy_counter[xy] / total
        p_y = y_counter[xy[1]] / total
        h += p_xy * np.log(p_y / p_xy)
    return h


def theils_u(x, y, symmetrical=False):
    """virus_grouping_estimator.py:39-71."""
    from collections import Counter
    from scipy.stats import entropy

    def marginal(v):
        c = Counter(v)
        tot = float(sum(c.values()))
        return entropy(list(map(lambda k: k / tot, c.values())))

    s_xy = _conditional_entropy(x, y)
    s_x = marginal(x)
    if not symmetrical:
        return 1.0 if s_x == 0 else (s_x - s_xy) / float(s_x)
    s_yx = _conditional_entropy(y, x)
    s_y = marginal(y)
    return 1.0 if s_x + s_y == 0.0 else ((s_x - s_xy) + (s_y - s_yx)) / float(s_x + s_y)


def VirusGrouping_Estimator(DistMat, Dendrogram_LinkageMethod, TaxoGroupingList, engine: Optional[Engine] = None):
    """app/utils/virus_grouping_estimator.py:74-89: the O(N^2) clustering runs on the device (bit-identical linkage matrix),
    the scalar search over the cut height and Theil's U stay on the host as in the reference (scipy)."""
    from scipy.cluster.hierarchy import fcluster
    from scipy.optimize import minimize_scalar
    _check_linkage_method(Dendrogram_LinkageMethod)
    linkageMat = (engine or default_engine()).linkage(np.asarray(DistMat, dtype=np.float64), Dendrogram_LinkageMethod)

    def anti_u(cut):
        return -1 * theils_u(x=fcluster(Z=linkageMat, t=cut, criterion="distance"), y=TaxoGroupingList, symmetrical=True)

    res = minimize_scalar(anti_u, bounds=[0, 1], method="bounded")
    groups = np.array(fcluster(Z=linkageMat, t=res.x, criterion="distance"))
    return (groups, res.x, -1 * res.fun, theils_u(x=TaxoGroupingList, y=groups), theils_u(x=groups, y=TaxoGroupingList))


# ------------------------------------------------------------- classification epilogue (pipeline II)
from .classification import (  # noqa: E402,F401
    PairwiseSimilarityScore_Cutoff_Dict_Constructor, max_similarity, draw_class_cells, cutoff_dict_from_cells, DeviceMatrix,
)


def bootstrap_classification(PPHMMSignatureTable_AllVirus, PPHMMLocationTable_AllVirus, TaxoGroupingList_RefVirus, GOMIDList,
                             payload, N_UcfViruses: int, N_Bootstrap: Optional[int] = None, indices=None, batch: int = 64,
                             engine: Optional[Engine] = None):
    """The numeric part of pipeline II's bootstrap loop body (app/src/virus_classification.py:290-344) with the similarity
    matrix resident in HBM: per replicate yields ``(cutoff_dict, MaxSimScoreList, TaxoOfMaxSimScoreList)`` -- what
    PairwiseSimilarityScore_Cutoff_Dict_Constructor(S[:N_ref, :N_ref], ...) and the 1-NN lines of
    TaxonomicAssignmentProposerAndEvaluator(S[-N_ucf:, :N_ref], ...) return -- reading only the listed cells and the row
    maxima from the device.  Every draw from numpy's global RNG happens in the reference's order (replicate b: resample
    indices, then per class the list subsamples and the SVC seed), which does not depend on device results."""
    from .engine import Job
    eng = engine or default_engine()
    sig = np.ascontiguousarray(PPHMMSignatureTable_AllVirus, dtype=np.float64)
    loc = np.ascontiguousarray(PPHMMLocationTable_AllVirus, dtype=np.float64)
    n, p = sig.shape
    taxo_ref = np.asarray(TaxoGroupingList_RefVirus)
    n_ref = len(taxo_ref)
    nb = int(payload["N_Bootstrap"] if N_Bootstrap is None else N_Bootstrap)
    n_pair = int(payload["N_PairwiseSimilarityScores"])
    scheme = payload["SimilarityMeasurementScheme"]
    if payload.get("PphmmNeighbourhoodWeight", 0) != 0 or payload.get("PphmmSigScoreThreshold", 0) != 0:
        raise NotImplementedError("batched bootstrap is built for PphmmNeighbourhoodWeight == 0 and threshold == 0")
    gor = gom_membership(taxo_ref, GOMIDList) if "G" in scheme else None
    job = Job(eng, sig, loc, gor, len(GOMIDList), scheme, float(payload.get("p", 1.0)), distance=False)
    try:
        for b0 in range(0, nb, batch):
            cb = min(batch, nb - b0)
            ws, cells = [], []
            for k in range(cb):             # the reference's draw order, replicate by replicate
                idx = draw_resample_indices(p, sort=True) if indices is None else indices[b0 + k]
                ws.append(weights_from_indices(idx, p))
                cells.append(draw_class_cells(taxo_ref, n_pair))
            out = []

            def consume(b, S):
                mx, tx = max_similarity(S, taxo_ref, row_begin=n - N_UcfViruses, n_ref=n_ref)
                out.append((cutoff_dict_from_cells(S, cells[b]), mx, tx))

            job.for_each_replicate_device(np.stack(ws), consume)
            for item in out:
                yield item
    finally:
        job.close()


def sort_pphmm_distances(PPHMMSignatureTable, engine: Optional[Engine] = None):
    """The matrix RefVirusAnnotator.sort_pphmm hands to DistMat2Tree (app/src/ref_virus_annotator.py:143-165):
    ``1 - SimMat_traits`` of the 0/1 transposed signature table, (P, P) float64, NaN where both PPHMMs are empty."""
    trait = np.ascontiguousarray(np.asarray(PPHMMSignatureTable, dtype=np.float64).T)
    return 1 - binary_jaccard(trait, engine)


def sort_pphmm_tree(PPHMMSignatureTable, do_logscale=False, engine: Optional[Engine] = None):
    """Newick string of the PPHMM dendrogram of sort_pphmm (:162-165): average linkage, leaves named 0..P-1.  The
    caller ladderizes it and reads the leaf order as the reference does (:167-170, Bio.Phylo)."""
    D = sort_pphmm_distances(PPHMMSignatureTable, engine)
    if not np.isfinite(D).all():
        # scipy's linkage refuses non-finite distances, so the reference fails here too (empty PPHMM columns are
        # removed upstream, remove_singleton_pphmm)
        raise ValueError("The condensed distance matrix must contain only finite values.")
    return DistMat2Tree(D, list(range(D.shape[0])), "average", do_logscale, engine=engine)


# ------------------------------------------------------------------------------ dendrograms
LINKAGE_METHODS = ("average", "complete", "weighted", "ward", "single", "centroid", "median")


def _check_linkage_method(method):
    """Every method scipy's ``linkage`` knows is built (api_classes.py:123 allows exactly these); anything else fails as
    scipy fails it."""
    if method not in LINKAGE_METHODS:
        raise ValueError(f"Invalid method: {method}")


def DistMat2Tree(DistMat, LeafList, Dendrogram_LinkageMethod, do_logscale=False, engine: Optional[Engine] = None):
    """app/utils/dist_mat_to_tree.py:5-21: Newick string of the UPGMA dendrogram of a distance matrix.  The clustering
    (scipy ``linkage`` of the upper triangle) runs on the device, to_tree + GetNewick natively on the host.  Like the
    reference, negative entries of the caller's matrix are zeroed in place.  Every method of scipy's ``linkage`` is built:
    the nearest-neighbour-chain methods (average -- the default of every shipped payload --, complete, weighted, ward), single
    (minimum spanning tree) and centroid / median (the heap-based generic algorithm; trees of at most 12 000 leaves)."""
    _check_linkage_method(Dendrogram_LinkageMethod)
    eng = engine or default_engine()
    DistMat[DistMat < 0] = 0
    D = np.asarray(DistMat, dtype=np.float64)
    if do_logscale:
        D = 10 ** D
    return eng.newick(eng.linkage(D, Dendrogram_LinkageMethod), LeafList)


def bootstrap_newick(PPHMMSignatureTable, PPHMMLocationTable, TaxoGroupingList, GOMIDList, payload, LeafList,
                     N_Bootstrap: Optional[int] = None, pipeline: str = "pl1", indices=None, batch: int = 512,
                     engine: Optional[Engine] = None) -> Iterator[str]:
    """One Newick line per bootstrap replicate, in the reference's order: the loop bodies of
    app/src/pl1_graphs.py:78-117 (resample, GOM rebuild, distance matrix, DistMat2Tree) with everything up to the
    dendrogram on the device -- the 8 N^2 bytes of a replicate's distance matrix never cross PCIe.  Arguments as
    ``bootstrap_distance_matrices``; ``LeafList`` are the tree labels (TaxoLabelList)."""
    sig = _as_table(PPHMMSignatureTable)
    n, p = sig.shape
    nb = int(payload["N_Bootstrap"] if N_Bootstrap is None else N_Bootstrap)
    scheme = payload["SimilarityMeasurementScheme"]
    method = payload.get("Dendrogram_LinkageMethod", "average")
    if payload.get("PphmmNeighbourhoodWeight", 0) != 0 or payload.get("PphmmSigScoreThreshold", 0) != 0:
        raise NotImplementedError("batched bootstrap is built for PphmmNeighbourhoodWeight == 0 and threshold == 0")
    _check_linkage_method(method)
    loc = sig if pipeline == "pl1" else _as_table(PPHMMLocationTable)
    gor = gom_membership(TaxoGroupingList, GOMIDList) if "G" in scheme else None
    job = _make_job(engine, sig, loc, gor, len(GOMIDList), scheme, float(payload.get("p", 1.0)), True)
    try:
        for b0 in range(0, nb, batch):
            cb = min(batch, nb - b0)
            if indices is None:
                idx = [draw_resample_indices(p, sort=(pipeline == "pl2")) for _ in range(cb)]
            else:
                idx = indices[b0:b0 + cb]
            w = np.stack([weights_from_indices(i, p) for i in idx])
            out = []
            job.for_each_tree(w, LeafList, lambda b, s: out.append(s), method)
            for s in out:
                yield s
    finally:
        job.close()


# ------------------------------------------------------------------------------ bootstrap
def _as_table(tab):
    """dense float64 C-order, or a scipy sparse matrix as it is (the *_coo tables of
    RefVirusAnnotator.p go to the device as triplets)"""
    from .engine import _is_sparse
    return tab if _is_sparse(tab) else np.ascontiguousarray(tab, dtype=np.float64)


def load_annotator_tables(pickle_file_or_dict, sparse: bool = True) -> dict:
    """The tables of RefVirusAnnotator.p / UcfVirusAnnotator.p (app/src/ref_virus_annotator.py:390-402,
    app/src/ucf_virus_annotator.py:53-58) in the form the batched entry points take: ``PPHMMSignatureTable`` and
    ``PPHMMLocationTable`` as scipy COO matrices when the pickle carries them (sparse=True), else dense; every other key is
    passed through."""
    import pickle
    d = pickle_file_or_dict
    if not isinstance(d, dict):
        with open(d, "rb") as fh:
            d = pickle.load(fh)
    out = dict(d)
    for key in ("PPHMMSignatureTable", "PPHMMLocationTable"):
        coo = d.get(key + "_coo")
        if coo is not None:
            from scipy.sparse import coo_matrix, issparse
            # the unclassified-virus pickle stores `coo_matrix(T).toarray()` under the _coo key (ucf_virus_annotator.py:56-57)
            out[key] = (coo if issparse(coo) else coo_matrix(np.asarray(coo))) if sparse else (coo.toarray() if issparse(coo) else np.asarray(coo))
    return out


def _make_job(engine, sig, loc, gor, g, scheme, p, distance):
    """One GPU (engine.Job) or every GPU of the node (multi.MultiJob): an explicit MultiEngine, or GRAVITY_B200_DEVICES
    naming more than one device when no engine is given."""
    from .engine import Job
    from .multi import MultiEngine, MultiJob, default_multi_engine
    if engine is None:
        engine = default_multi_engine()
    if isinstance(engine, MultiEngine):
        from .engine import _is_sparse
        dense = lambda a: a.toarray() if _is_sparse(a) else a  # noqa: E731  (the multi-GPU entry takes dense host tables)
        return MultiJob(engine, dense(sig), sig if loc is sig else dense(loc), gor, g, scheme, p, distance=distance)
    return Job(engine or default_engine(), sig, loc, gor, g, scheme, p, distance=distance)


def draw_resample_indices(n_pphmms: int, sort: bool = False):
    """The reference's own draw from numpy's global legacy RNG (pl1_graphs.py:81-82; sorted in
    virus_classification.py:295-296), keeping that stream in step with the reference's later draws."""
    idx = np.random.choice(list(range(n_pphmms)), n_pphmms, replace=True)
    return np.array(sorted(idx)) if sort else idx


def weights_from_indices(idx, n_pphmms: int):
    return np.bincount(np.asarray(idx, dtype=np.int64), minlength=n_pphmms).astype(np.int32)


def gom_membership(TaxoGroupingList, GOMIDList, n_rows: Optional[int] = None):
    """int32 GOM index per table row (-1 = not in any GOM), the device form of GOMDB_Constructor."""
    taxo = np.asarray(TaxoGroupingList)
    out = np.full(len(taxo) if n_rows is None else n_rows, -1, dtype=np.int32)
    for k, g in enumerate(GOMIDList):
        out[: len(taxo)][taxo == g] = k
    return out


def bootstrap_distance_matrices(PPHMMSignatureTable, PPHMMLocationTable, TaxoGroupingList, GOMIDList, payload,
                                N_Bootstrap: Optional[int] = None, pipeline: str = "pl1", indices=None,
                                batch: int = 256, distance: bool = True,
                                engine: Optional[Engine] = None, condensed: bool = False) -> Iterator[np.ndarray]:
    """Yields one (N, N) float64 matrix per bootstrap replicate, in the reference's order.
    condensed=True (one GPU): yields scipy's condensed form instead -- the N (N - 1) / 2 entries above the diagonal, what
    dist_mat_to_tree.py:8-9 computes before linkage() -- so half the bytes cross PCIe.

    pipeline="pl1": app/src/pl1_graphs.py:78-107 -- unsorted indices, and the bootstrapped SIGNATURE
    table doubles as the location table; yields D.
    pipeline="pl2": app/src/virus_classification.py:290-315 -- sorted indices, real location table,
    GOMDB from the first len(TaxoGroupingList) rows.
    ``indices`` (B, P) overrides the draw (e.g. Engine.philox_indices); by default indices are drawn
    exactly as the reference draws them, replicate by replicate.
    Each yielded matrix is a zero-copy view into a pinned host ring: it is valid until the next one is
    requested (copy it to keep it), which is how the reference's loop uses it (tree, then discard)."""
    sig = _as_table(PPHMMSignatureTable)
    n, p = sig.shape
    nb = int(payload["N_Bootstrap"] if N_Bootstrap is None else N_Bootstrap)
    scheme = payload["SimilarityMeasurementScheme"]
    if payload.get("PphmmNeighbourhoodWeight", 0) != 0 or payload.get("PphmmSigScoreThreshold", 0) != 0:
        raise NotImplementedError("batched bootstrap is built for PphmmNeighbourhoodWeight == 0 and threshold == 0")
    loc = sig if pipeline == "pl1" else _as_table(PPHMMLocationTable)
    gor = gom_membership(TaxoGroupingList, GOMIDList) if "G" in scheme else None
    job = _make_job(engine, sig, loc, gor, len(GOMIDList), scheme, float(payload.get("p", 1.0)), distance)
    if condensed and not hasattr(job, "ntiles"):
        job.close()
        raise NotImplementedError("condensed delivery is built for the single-GPU engine")
    try:
        # tables are uploaded once; replicates go through in batches big enough to fill the tensor-core
        # blocks (128 replicates per CTA) and the 32-replicate lanes of the GOM kernels
        for b0 in range(0, nb, batch):
            cb = min(batch, nb - b0)
            if indices is None:
                idx = [draw_resample_indices(p, sort=(pipeline == "pl2")) for _ in range(cb)]
            else:
                idx = indices[b0:b0 + cb]
            w = np.stack([weights_from_indices(i, p) for i in idx])
            for m in (job.replicates(w, condensed=True) if condensed else job.replicates(w)):
                yield m
    finally:
        job.close()
