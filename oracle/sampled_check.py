"""Sampled-cell parity checks at full BASELINE sizes (TEST INFRASTRUCTURE: imported only by tests/, by
__graft_entry__.smoke() and by bench.py's parity_check / cpu_baseline legs -- never by the product).

At the benched shapes (10 000 x 30 000 and 25 000 x 30 000 tables, 200 GOMs) a whole replicate matrix costs the
CPU oracle hours, but single cells are cheap: one (replicate, genome, GOM) cell of the GOM signature table is one
``weighted_dcor`` over ~2 000 points, one (replicate, i, j) cell of the distance matrix is two row reductions.
The functions below restate the reference's expressions for exactly one cell:

  gom_cell   app/utils/parallel_gom_sig_generator.py:30-37 on the resampled tables of
             app/src/pl1_graphs.py:83-100 (PL1: the bootstrapped signature table doubles as location table,
             quirk Q3) or app/src/virus_classification.py:297-308 (PL2: real location table, GOMDB from the
             reference rows only);
  pair_cell  app/utils/similarity_matrix_constructor.py:139-140,167-177,201-226 and app/src/pl1_graphs.py:52-53.

A column resample ``tab[:, idx]`` is evaluated through its multiplicities ``w = bincount(idx)``; ``gom_cell``
repeats the points literally (oracle.weighted_dcor), so the device's multiplicity algebra is checked
independently.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import gravity_oracle as O


def gom_cell(loc: np.ndarray, member_rows: np.ndarray, v: int, w: np.ndarray | None = None) -> float:
    """GOM signature of genome ``v`` against the GOM whose members are rows ``member_rows`` of ``loc``, for the
    column multiplicities ``w`` (None = the un-resampled table)."""
    gom_rows = loc[np.asarray(member_rows, dtype=np.int64)]
    y = loc[v]
    ww = np.ones(loc.shape[1], dtype=np.int64) if w is None else np.asarray(w, dtype=np.int64)
    in_gom = (gom_rows != 0).any(axis=0) if len(gom_rows) else np.zeros(loc.shape[1], bool)
    rel = np.where((in_gom | (y != 0)) & (ww > 0))[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.weighted_dcor(gom_rows[:, rel].T, y[rel].reshape(-1, 1), ww[rel])


def _gj_rows(a: np.ndarray, b: np.ndarray, w: np.ndarray | None) -> float:
    mn, mx = np.minimum(a, b), np.maximum(a, b)
    if w is not None:
        mn, mx = mn * w, mx * w
    smax = np.sum(mx)
    with np.errstate(invalid="ignore", divide="ignore"):
        v = np.sum(mn) / smax if not smax == 0 else 0.0          # similarity_matrix_constructor.py:167-168
    if np.isnan(v) or v < 0:                                      # :201-205
        v = 0.0
    return float(v)


def pair_cell(sig: np.ndarray | None, gom_i: np.ndarray | None, gom_j: np.ndarray | None, i: int, j: int,
              w: np.ndarray | None = None, scheme: str = "PG", p: float = 1.0, distance: bool = True) -> float:
    """One entry of the (replicate) similarity / distance matrix.  ``gom_i`` / ``gom_j`` are the two genomes' rows of
    the replicate's GOM signature table; ``w`` the PPHMM column multiplicities of the replicate."""
    vp = vg = 0.0
    if "P" in scheme:
        a, b = np.array(sig[i], dtype=np.float64), np.array(sig[j], dtype=np.float64)
        a[a < 0] = 0                                              # `row[row < thr] = 0` with thr == 0, :139-140
        b[b < 0] = 0
        ww = None if w is None else np.asarray(w, dtype=np.float64)
        if ww is not None:                                        # a NaN in a column that was not drawn is not in the resample
            a, b = np.where(ww > 0, a, 0.0), np.where(ww > 0, b, 0.0)
        vp = _gj_rows(a, b, ww)
    if "G" in scheme:
        vg = _gj_rows(np.asarray(gom_i, dtype=np.float64), np.asarray(gom_j, dtype=np.float64), None)
    s = {"P": vp, "G": vg, "PG": (vp * vg) ** 0.5}[scheme]
    s = max(s, 0.0) ** float(p)                                   # :224-226
    return max(1.0 - s, 0.0) if distance else s                   # pl1_graphs.py:52-53


def sample_pairs(rng: np.random.Generator, n: int, taxon_of: np.ndarray, k: int, special=()):
    """k (i, j) pairs: the given special ones, a few diagonal / same-taxon pairs, the rest uniform."""
    out = list(special)
    tax = np.asarray(taxon_of)
    for _ in range(max(2, k // 10)):
        i = int(rng.integers(0, n))
        same = np.where(tax == tax[i])[0]
        out.append((i, int(same[rng.integers(0, len(same))])))
    out.append((int(rng.integers(0, n)),) * 2)
    while len(out) < k:
        out.append((int(rng.integers(0, n)), int(rng.integers(0, n))))
    return out[:k] if len(out) > k else out
