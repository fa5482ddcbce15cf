"""CPU oracle for the GRAViTy-V2 pairwise genome-similarity hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gravity2_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the timed CPU baseline.

Every function restates one piece of the reference (Mayne941/gravity2, all file:line
relative to the reference checkout) in plain numpy, keeping the reference's own
expressions wherever they decide the floating-point result, so that the restatement is
bit-identical to the imported reference on the same inputs.  Parity is PINNED: the
restatement is checked against
  * the reference's own known answers (dcor docstring KAT ``app/utils/dcor.py:38-41``,
    ``test/utils/test_dcor.py:32-34`` cent_dist literals, ``test_gomdb_constructor.py:6-28``,
    ``test_dist_mat_to_tree.py:17`` Newick golden), and
  * outputs of the UNMODIFIED reference functions imported from /root/reference in the
    build container, committed as ``tests/golden/*.npz`` by ``oracle/make_golden.py``.
See ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "cent_dist", "dcov", "dvar", "dcor", "gomdb", "generate_gom_sigs", "gom_table",
    "similarity_mat", "similarity_mat_fast", "neighbourhood_index_lists", "gj_pair_sums", "dist_from_sim",
    "shared_pphmm_ratio", "shared_norm_pphmm_ratio", "binary_gj", "bootstrap_indices",
    "pl1_bootstrap_replicate", "pl2_bootstrap_replicate", "weights_from_indices",
    "weighted_gj", "weighted_dcor", "dist_mat_to_tree", "linkage_merges",
]


# ----------------------------------------------------------------------------- dcor
def _pdist_square(X: np.ndarray) -> np.ndarray:
    """``squareform(pdist(X))`` (Euclidean), app/utils/dcor.py:24 -- the same scipy call
    the reference makes (scipy is a third-party dependency of the reference, pinned
    scipy==1.14.0 in requirements.txt:4; this image has 1.18.1)."""
    from scipy.spatial.distance import pdist, squareform
    return squareform(pdist(np.asarray(X, dtype=np.float64)))


def cent_dist(X: np.ndarray) -> np.ndarray:
    """Double-centred distance matrix, app/utils/dcor.py:19-32."""
    M = _pdist_square(X)
    rmean = M.mean(axis=1)
    cmean = M.mean(axis=0)
    gmean = rmean.mean()
    R = np.tile(rmean, (M.shape[0], 1)).transpose()
    C = np.tile(cmean, (M.shape[1], 1))
    G = np.tile(gmean, M.shape)
    return M - R - C + G


def dcov(X: np.ndarray, Y: np.ndarray) -> float:
    """app/utils/dcor.py:4-11 (sqrt of a possibly tiny-negative sum -> NaN, quirk Q5)."""
    n = X.shape[0]
    with np.errstate(invalid="ignore"):
        return np.sqrt(np.multiply(X, Y).sum()) / n


def dvar(X: np.ndarray) -> float:
    """app/utils/dcor.py:13-17."""
    return np.sqrt(np.sum(X ** 2 / X.shape[0] ** 2))


def dcor(X: np.ndarray, Y: np.ndarray) -> float:
    """Distance correlation, app/utils/dcor.py:34-62.  KAT: dcor([1..5],[1,2,9,4,4]) =
    0.7626762424168665."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    assert X.shape[0] == Y.shape[0]
    if X.shape[0] == 0:          # reference: empty pdist -> dvar of empty = 0.0 -> 0.0 (Q6)
        return 0.0
    A = cent_dist(X)
    B = cent_dist(Y)
    dcov_AB = dcov(A, B)
    dvar_A = dvar(A)
    dvar_B = dvar(B)
    out = 0.0
    if dvar_A > 0.0 and dvar_B > 0.0:
        out = dcov_AB / np.sqrt(dvar_A * dvar_B)
    return float(out)


def weighted_dcor(X: np.ndarray, Y: np.ndarray, w: np.ndarray) -> float:
    """dcor of the point set in which point k is repeated w[k] times (the bootstrap
    column resample seen from the GOM kernel's side).  Computed
    literally, by repeating the rows, so that it is an independent check of the
    multiplicity algebra used on the device."""
    idx = np.repeat(np.arange(len(w)), np.asarray(w, dtype=np.int64))
    return dcor(np.asarray(X)[idx], np.asarray(Y)[idx])


# ------------------------------------------------------------------------------ GOM
def gomdb(taxo_grouping_list, loc_table, gom_id_list) -> dict:
    """GOMDB_Constructor, app/utils/gomdb_constructor.py:2-8."""
    taxo = np.asarray(taxo_grouping_list)
    return {g: loc_table[taxo == g, :] for g in gom_id_list}


def generate_gom_sigs(gom_rows: np.ndarray, loc_table: np.ndarray) -> list:
    """One GOM column, app/utils/parallel_gom_sig_generator.py:30-37.

    The reference's ``map(any, zip(map(any, GOM.T != 0), loc != 0))`` is the boolean
    ``(GOM != 0).any(axis=0) | (loc != 0)``."""
    gom_rows = np.asarray(gom_rows, dtype=np.float64)
    in_gom = (gom_rows != 0).any(axis=0) if gom_rows.shape[0] else np.zeros(gom_rows.shape[1], bool)
    out = []
    for loc in np.asarray(loc_table, dtype=np.float64):
        rel = np.where(in_gom | (loc != 0))[0]
        out.append(dcor(gom_rows[:, rel].T, loc[rel].reshape(-1, 1)))
    return out


def gom_table(loc_table, gom_db: dict, gom_id_list) -> np.ndarray:
    """GOMSignatureTable_Constructor, parallel_gom_sig_generator.py:10-28 (Pool removed;
    column order = gom_id_list, :24-26)."""
    N = len(loc_table)
    tab = np.empty((N, 0))
    for g in gom_id_list:
        tab = np.column_stack((tab, generate_gom_sigs(gom_db[g], loc_table)))
    return tab


# ----------------------------------------------------------------------- similarity
def _gj(a: np.ndarray, b: np.ndarray) -> float:
    """similarity_matrix_constructor.py:167-168 (the reference's literal expression)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        mx = np.sum(np.maximum(a, b))
        return np.sum(np.minimum(a, b)) / mx if not mx == 0 else 0


def neighbourhood_index_lists(loc_csv: np.ndarray, sig_csv: np.ndarray):
    """The CSV preprocessing of SimilarityMat_Constructor,
    app/utils/similarity_matrix_constructor.py:19-110, restated loop for loop.

    ``loc_csv``: numeric block of ``fnames["PphmmLocs"]`` (``iloc[:, 1:]``, :20-22);
    ``sig_csv``: numeric block of ``fnames["PphmmAndGomSigs"]`` (``iloc[:, 2:]``, :29-31).
    Returns ``(neg, pos)``: dict genome row -> list of ORIGINAL column numbers whose entry is
    multiplied by (1 - w) / (1 + w) every time the pair loop touches that row (:120-133)."""
    import pandas as pd
    trim = pd.DataFrame(np.asarray(loc_csv, dtype=np.float64)).replace(0, np.nan)      # :24
    means = trim.mean().to_list()                                                      # :25
    means_indices = np.argsort(means)                                                  # :26
    sig_nan = np.array(pd.DataFrame(np.asarray(sig_csv, dtype=np.float64)).replace(0, np.nan))   # :32
    all_dists = []
    for row in sig_nan:                                                                # :35-39
        dists = []
        for i in range(row.shape[0]):
            dists.append(round(np.abs(row[i] - means[i]), 4) if not row[i] == np.nan else 0)
        all_dists.append(dists)
    distance_arr = np.nan_to_num(np.array(all_dists).T[means_indices].T, 0)            # :41
    neighbourhoods = {}
    for row_idx, row in enumerate(distance_arr):                                       # :45-77 (AMP_SINGLETONS = True)
        neighbourhood = 0
        for prof_idx, prof in enumerate(row):
            if prof_idx == 0:
                continue
            if prof != 0:
                neighbourhood += 1
                if prof_idx == len(row) - 1:
                    neighbourhoods.setdefault(row_idx, {}).update(
                        {prof_idx - neighbourhood: [neighbourhood + 1, row[prof_idx - neighbourhood:prof_idx]]})
            else:
                if neighbourhood == 0:
                    continue
                neighbourhoods.setdefault(row_idx, {}).update(
                    {prof_idx - neighbourhood: [neighbourhood + 1, row[prof_idx - neighbourhood - 1:prof_idx]]})
                neighbourhood = 0
    neg_sorted, pos_sorted = {}, {}
    for row_key, row in neighbourhoods.items():                                        # :79-95 (MIN_NEIGH_SIZE = 5)
        for nei_key, nei in row.items():
            if nei[0] < 5:
                continue
            neg_sorted.setdefault(row_key, [])
            idxs = [i + nei_key for i in range(len(nei[1]))]
            pos_sorted.setdefault(row_key, [])
            pos_sorted[row_key] = pos_sorted[row_key] + [np.argmax(nei[1]) + nei_key]
            del idxs[np.argmax(nei[1])]
            neg_sorted[row_key] = neg_sorted[row_key] + idxs
    reversed_ = np.argsort(means_indices)                                              # :98-110
    neg = {r: [int(np.where(reversed_ == prof)[0][0]) for prof in row] for r, row in neg_sorted.items()}
    pos = {r: [int(np.where(reversed_ == prof)[0][0]) for prof in row] for r, row in pos_sorted.items()}
    return neg, pos


def similarity_mat(sig, gom, loc, spr, neighbourhood_weight=0.0, score_threshold=0,
                   scheme="PG", p=1.0, neighbourhoods=None) -> np.ndarray:
    """SimilarityMat_Constructor pair loop + clean/combine/power,
    app/utils/similarity_matrix_constructor.py:112-226.

    With ``neighbourhood_weight == 0`` (the default, api_classes.py:135, and every shipped
    payload) the CSV preprocessing :19-110 has no effect on the result (quirk Q2) and the
    in-place row mutation :126-140 reduces to the idempotent threshold ``row[row < thr] = 0``
    (quirk Q1).  With a non-zero weight pass ``neighbourhoods = neighbourhood_index_lists(..)``:
    the rows are then scaled IN PLACE every time the pair loop touches them, exactly as the
    reference does (the caller's ``sig`` is left mutated, and the result depends on it)."""
    if neighbourhood_weight != 0 and neighbourhoods is None:
        raise ValueError("neighbourhood_weight != 0 needs the index lists of neighbourhood_index_lists()")
    neg, pos = neighbourhoods if neighbourhoods is not None else ({}, {})
    N = sig.shape[0] if sig is not None else gom.shape[0]
    p = float(p)
    P_mat = np.zeros((N, N))
    G_mat = np.zeros((N, N))
    L_mat = np.zeros((N, N))
    NEG_WEIGHT = 1 - neighbourhood_weight
    POS_WEIGHT = 1 + neighbourhood_weight
    for i in range(N):
        for j in range(i, N):
            if "P" in scheme:
                si = sig[i]
                sj = sig[j]
                for r, row in ((i, si), (j, sj)):                                      # :126-137
                    if r in neg:
                        row[neg[r]] = row[neg[r]] * NEG_WEIGHT
                        row[pos[r]] = row[pos[r]] * POS_WEIGHT
                si[si < score_threshold] = 0
                sj[sj < score_threshold] = 0
                P_mat[i, j] = _gj(si, sj)
                P_mat[j, i] = P_mat[i, j]
            if "G" in scheme:
                G_mat[i, j] = _gj(gom[i], gom[j])
                G_mat[j, i] = G_mat[i, j]
            if "L" in scheme:
                li, lj = loc[i], loc[j]
                present = np.column_stack((li != 0, lj != 0)).any(axis=1)
                L_mat[i, j] = dcor(li[present].reshape(-1, 1), lj[present].reshape(-1, 1))
                L_mat[j, i] = L_mat[i, j]
    return _combine(P_mat, G_mat, L_mat, spr, scheme, p)


def _combine(P_mat, G_mat, L_mat, spr, scheme, p):
    """similarity_matrix_constructor.py:196-226."""
    for k, v in {"P": P_mat, "G": G_mat, "L": L_mat}.items():
        if k in scheme:
            v[np.where(np.isnan(v))] = 0
            v[v < 0] = 0
    if scheme == "P":
        S = P_mat
    elif scheme == "G":
        S = G_mat
    elif scheme == "L":
        S = L_mat
    elif scheme == "PG":
        S = (P_mat * G_mat) ** 0.5
    elif scheme == "PL":
        S = P_mat * L_mat
    elif scheme == "R":
        S = spr
    elif scheme == "RG":
        S = (spr * G_mat) ** 0.5
    elif scheme == "PR":
        S = (spr * P_mat) ** 0.5
    else:
        raise ValueError(scheme)
    S[S < 0] = 0
    return S ** p


def gj_pair_sums(tab: np.ndarray, w: np.ndarray | None = None):
    """(sum_min, sum_max) for all pairs in float64, row-block vectorised; ``w`` = integer
    column multiplicities (bootstrap).  Used by ``similarity_mat_fast``."""
    tab = np.asarray(tab, dtype=np.float64)
    N = tab.shape[0]
    ww = None if w is None else np.asarray(w, dtype=np.float64)
    smin = np.zeros((N, N))
    smax = np.zeros((N, N))
    for i in range(N):
        mn = np.minimum(tab[i], tab[i:])
        mx = np.maximum(tab[i], tab[i:])
        if ww is not None:
            mn = mn * ww
            mx = mx * ww
        smin[i, i:] = mn.sum(axis=1)
        smax[i, i:] = mx.sum(axis=1)
        smin[i:, i] = smin[i, i:]
        smax[i:, i] = smax[i, i:]
    return smin, smax


def weighted_gj(tab, w=None) -> np.ndarray:
    """GJ matrix with column multiplicities; equals ``_gj`` on ``tab[:, idx]`` up to
    summation order."""
    smin, smax = gj_pair_sums(tab, w)
    with np.errstate(invalid="ignore", divide="ignore"):
        g = smin / smax
    # `if not sum(max) == 0 else 0`; NaN != 0 so NaN/NaN stays NaN like the reference
    g[smax == 0] = 0
    return g


def similarity_mat_fast(sig, gom, spr=None, score_threshold=0, scheme="PG", p=1.0,
                        w=None) -> np.ndarray:
    """Row-block vectorised equivalent of ``similarity_mat`` for schemes without "L"
    (agrees to summation-order rounding, ~1e-15; tests pin <=1e-12).  ``w`` applies the
    bootstrap multiplicities to the PPHMM signature columns."""
    assert "L" not in scheme
    N = sig.shape[0] if sig is not None else gom.shape[0]
    P_mat = np.zeros((N, N))
    G_mat = np.zeros((N, N))
    if "P" in scheme:
        s = np.array(sig, dtype=np.float64)
        s[s < score_threshold] = 0
        P_mat = weighted_gj(s, w)
    if "G" in scheme:
        G_mat = weighted_gj(gom)
    return _combine(P_mat, G_mat, np.zeros((N, N)), None if spr is None else np.array(spr), scheme, float(p))


def dist_from_sim(S: np.ndarray) -> np.ndarray:
    """app/src/pl1_graphs.py:52-53; virus_classification.py:158-159."""
    D = 1 - S
    D[D < 0] = 0
    return D


# ----------------------------------------------------------------- presence/absence
def shared_pphmm_ratio(arr: np.ndarray) -> list:
    """app/utils/shared_pphmm_graphs.py:59-77 on the already-loaded table: list of lists
    of Python ints, I[i][j] = #{c: arr[i,c] != 0 and arr[j,c] != 0}."""
    nz = (np.asarray(arr) != 0)
    out = []
    for row in nz:
        out.append([int(np.count_nonzero(row & other)) for other in nz])
    return out


def shared_norm_pphmm_ratio(arr: np.ndarray) -> np.ndarray:
    """app/utils/shared_pphmm_graphs.py:79-97.  ``min(hits, mp) / max(hits, mp)`` on Python
    ints: 0/0 raises ZeroDivisionError exactly as the reference does (:86)."""
    nz = (np.asarray(arr) != 0)
    cnt = nz.sum(axis=1)
    n = nz.shape[0]
    out = np.zeros((n, n))
    for ia in range(n):
        for ib in range(n):
            mp = int(min(cnt[ia], cnt[ib]))
            hits = int(np.count_nonzero(nz[ia] & nz[ib]))
            out[ib, ia] = min([hits, mp]) / max([hits, mp])
    return out


def binary_gj(trait01: np.ndarray) -> np.ndarray:
    """PPHMM x PPHMM binary Jaccard of RefVirusAnnotator.sort_pphmm,
    app/src/ref_virus_annotator.py:148-158 (no zero guard: empty-vs-empty -> NaN)."""
    T = np.array(trait01, dtype=np.float64)
    T[T != 0] = 1
    n = len(T)
    rows = []
    with np.errstate(invalid="ignore", divide="ignore"):
        for tv in T:
            M = np.tile(tv, (n, 1))
            rows.append(np.sum(np.minimum(M, T), axis=1) / np.sum(np.maximum(M, T), axis=1))
    return np.array(rows)


# ------------------------------------------------------------------------ bootstrap
def bootstrap_indices(n_pphmms: int, sort: bool = False) -> np.ndarray:
    """One replicate's resample indices from the GLOBAL legacy numpy RNG, exactly the
    reference's call (pl1_graphs.py:81-82; sorted in virus_classification.py:295-296).
    The caller seeds ``np.random.seed`` (the reference never does, quirk Q4)."""
    idx = np.random.choice(list(range(n_pphmms)), n_pphmms, replace=True)
    return np.array(sorted(idx)) if sort else idx


def weights_from_indices(idx: np.ndarray, n_pphmms: int) -> np.ndarray:
    return np.bincount(np.asarray(idx, dtype=np.int64), minlength=n_pphmms).astype(np.int32)


def pl1_bootstrap_replicate(sig, taxo, gom_ids, idx, scheme="PG", p=1.0, spr=None):
    """One pass of the PL1 loop body, app/src/pl1_graphs.py:83-107, incl. quirk Q3 (the
    bootstrapped *signature* table is used as the location table)."""
    b_sig = sig[:, idx]
    b_loc = sig[:, idx]
    b_gom = None
    if "G" in scheme:
        db = gomdb(taxo, b_loc, gom_ids)
        b_gom = gom_table(b_loc, db, gom_ids)
    S = similarity_mat(b_sig.copy(), b_gom, b_loc, spr, 0.0, 0, scheme, p)
    return dist_from_sim(S)


def pl2_bootstrap_replicate(sig_all, loc_all, n_ref, taxo_ref, gom_ids, idx, scheme="PG", p=1.0, spr=None):
    """One pass of the PL2 loop body, app/src/virus_classification.py:297-315 (sorted
    indices, real location table, GOMDB from the reference rows only).  Returns (S, D)."""
    b_sig = sig_all[:, idx]
    b_loc = loc_all[:, idx]
    b_gom = None
    if "G" in scheme:
        db = gomdb(taxo_ref, b_loc[:n_ref], gom_ids)
        b_gom = gom_table(b_loc, db, gom_ids)
    S = similarity_mat(b_sig.copy(), b_gom, b_loc, spr, 0.0, 0, scheme, p)
    return S, dist_from_sim(S.copy())


# ------------------------------------------------------------------ downstream check
def linkage_merges(D: np.ndarray, method: str = "average"):
    """scipy linkage on the condensed upper triangle, as dist_mat_to_tree.py:7-12 does;
    returned for the dendrogram-topology parity check."""
    from scipy.cluster.hierarchy import linkage
    D = np.array(D)
    D[D < 0] = 0
    return linkage(D[np.triu_indices_from(D, k=1)], method=method)


def _newick(node, newick, parentdist, leaf_names):
    """app/utils/get_newick.py:1-13."""
    if node.is_leaf():
        return "%s:%f%s" % (leaf_names[node.id], parentdist - node.dist, newick)
    if len(newick) > 0:
        newick = "):%f%s" % (parentdist - node.dist, newick)
    else:
        newick = ");"
    newick = _newick(node.get_left(), newick, node.dist, leaf_names)
    newick = _newick(node.get_right(), ",%s" % (newick), node.dist, leaf_names)
    return "(%s" % (newick)


def dist_mat_to_tree(D, leaf_list, method="average", do_logscale=False) -> str:
    """DistMat2Tree, app/utils/dist_mat_to_tree.py:5-21."""
    from scipy.cluster.hierarchy import linkage, to_tree
    D = np.array(D)
    D[D < 0] = 0
    dl = D[np.triu_indices_from(D, k=1)]
    if do_logscale:
        dl = 10 ** dl
    tree = to_tree(Z=linkage(dl, method=method), rd=False)
    return _newick(tree, "", tree.dist, leaf_list)
