"""Host side of the PphmmNeighbourhoodWeight != 0 mode of ``SimilarityMat_Constructor``.

The reference (app/utils/similarity_matrix_constructor.py:19-110) re-reads two CSVs on every
call, orders the PPHMMs by mean genomic location, finds each genome's *neighbourhoods* (runs of
at least four consecutive hits in that order) and from them a negative and a positive index list
per genome; the pair loop (:112-140) then rescales those entries of the caller's table in place
every time it touches a row.

Here the lists are built with array operations (run boundaries from one ``diff`` over the hit
mask) as a dense uint8 class table -- bit 0: column on the genome's negative list, bit 1: on its
positive list -- which is what ``gv2_similarity_nbhd_host`` takes, and the table mutation the
reference leaves behind (every row through N + 1 applications) is replayed on the flagged
entries only.
"""
from __future__ import annotations

import numpy as np

MIN_NEIGH_SIZE = 5          # similarity_matrix_constructor.py:79 (stored length = run length + 1)
NEG, POS = 1, 2


def read_neighbourhood_csvs(fnames):
    """The numeric blocks the reference reads: ``PphmmLocs`` without its first column (:20-23),
    ``PphmmAndGomSigs`` without its first two (:29-31)."""
    import pandas as pd
    loc = pd.read_csv(fnames["PphmmLocs"], index_col=False).iloc[:, 1:].apply(pd.to_numeric)
    sig = pd.read_csv(fnames["PphmmAndGomSigs"], index_col=False).iloc[:, 2:].apply(pd.to_numeric)
    return np.asarray(loc, dtype=np.float64), np.asarray(sig, dtype=np.float64)


def location_order(loc_csv: np.ndarray) -> np.ndarray:
    """``means_indices`` (:24-26): PPHMM columns by mean location over the genomes that hit them
    (zeros are missing values; a never-hit column has a NaN mean and sorts last)."""
    import pandas as pd
    # pandas' own reduction (skipna mean), so that near-tied means order exactly as in the reference
    means = np.asarray(pd.DataFrame(loc_csv).replace(0, np.nan).mean(), dtype=np.float64)
    return means, np.argsort(means.tolist())


def neighbourhood_classes(loc_csv: np.ndarray, sig_csv: np.ndarray) -> np.ndarray:
    """uint8 (N, P) class table: NEG / POS bits per (genome, original PPHMM column)."""
    n, p = sig_csv.shape
    means, order = location_order(loc_csv)
    with np.errstate(invalid="ignore"):
        dist = np.round(np.abs(np.where(sig_csv == 0, np.nan, sig_csv) - means[None, :]), 4)      # :35-39
    dist = np.nan_to_num(dist[:, order], nan=0.0)                                                  # :41
    hit = dist != 0
    cls_sorted = np.zeros((n, p), dtype=np.uint8)
    if p < 2:
        return cls_sorted
    # runs of hits among positions 1 .. p-1 (position 0 is skipped by the scan, :46-50)
    h = np.zeros((n, p + 1), dtype=np.int8)
    h[:, 1:p] = hit[:, 1:]
    d = np.diff(h, axis=1)                      # d[:, k] = h[k+1] - h[k]
    rows_s, pos_s = np.nonzero(d == 1)          # run starts at s = pos_s + 1
    rows_e, pos_e = np.nonzero(d == -1)         # run ends at e = pos_e (last hit)
    # starts and ends pair up in order within each row (np.nonzero is row-major)
    starts, ends = pos_s + 1, pos_e
    length = ends - starts + 1
    keep = length + 1 >= MIN_NEIGH_SIZE                                                            # :81
    for r, s, e in zip(rows_s[keep], starts[keep], ends[keep]):
        if e == p - 1:      # run reaches the last column: values row[s-1 : p-1], candidates s-1 .. p-2   (:59-65)
            first, vals = s - 1, dist[r, s - 1:p - 1]
        else:               # zero-terminated: values row[s-1 : e+1], candidates s .. e+1                  (:67-77)
            first, vals = s, dist[r, s - 1:e + 1]
        k = int(np.argmax(vals))                                                                   # :90-93
        cand = np.arange(first, first + len(vals))
        cls_sorted[r, cand[k]] |= POS
        cls_sorted[r, np.delete(cand, k)] |= NEG
    cls = np.zeros_like(cls_sorted)
    cls[:, order] = cls_sorted                  # sorted position q is original column order[q]   (:97-110)
    return cls


def apply_in_place(table: np.ndarray, cls: np.ndarray, weight: float, threshold: float, applications: int) -> None:
    """Leave the caller's table as the reference leaves it: flagged entries through ``applications``
    rounds of ``x * (1 - w)`` / ``x * (1 + w)`` / ``x < thr -> 0`` (:126-140), the rest thresholded."""
    neg_w, pos_w = 1 - weight, 1 + weight
    rr, cc = np.nonzero(cls)
    k = cls[rr, cc]
    v = table[rr, cc].astype(np.float64)
    is_neg, is_pos = (k & NEG) != 0, (k & POS) != 0
    for _ in range(applications):
        v = np.where(is_neg, v * neg_w, v)
        v = np.where(is_pos, v * pos_w, v)
        v = np.where(v < threshold, 0.0, v)
    table[table < threshold] = 0
    table[rr, cc] = v
