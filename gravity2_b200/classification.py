"""Classification epilogue of pipeline II: the similarity cut-off per taxonomic class and the
1-nearest-neighbour proposal, without the reference's O(N^2) Python loop and -- when the similarity matrix lives in HBM --
without shipping the matrix.

Reference (Mayne941/gravity2):
  PairwiseSimilarityScore_Cutoff_Dict_Constructor   app/utils/classification_utils.py:10-60
  TaxonomicAssignmentProposerAndEvaluator (1-NN part) app/utils/classification_utils.py:63-68

The reference walks every pair (i, j >= i) and appends SimMat[i][j] to per-class intra / inter lists, then subsamples a
list longer than N_PairwiseSimilarityScores with ``np.random.choice(list, N, replace=False)`` and fits a linear SVC.  The
ORDER of a list matters (the subsample picks by position), but it does not depend on the matrix: here the positions are
computed in closed form from the class membership, the legacy global RNG is consumed exactly as the reference consumes
it (``choice`` without replacement draws ``permutation(len(list))[:N]`` whatever the list holds), and only the chosen
cells are read -- by numpy indexing for a host matrix, by ``gv2_dev_gather_cells`` for a device-resident one.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .engine import Engine, default_engine


class DeviceMatrix:
    """A float64 (n, n) matrix in HBM (a slot of a job's device ring): pointer + leading dimension, valid only while the
    callback that received it runs."""

    def __init__(self, eng: Engine, ptr: int, n: int, ld: Optional[int] = None):
        self.eng, self.ptr, self.n, self.ld = eng, int(ptr), int(n), int(n if ld is None else ld)
        self.shape = (self.n, self.n)

    def __len__(self):
        return self.n

    def gather(self, rows: np.ndarray, cols: np.ndarray) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        k = len(rows)
        out = np.empty(k, dtype=np.float64)
        if k == 0:
            return out
        lib, ctx = self.eng.lib, self.eng.ctx
        d = C.c_void_p()
        _lib.check(ctx, lib.gv2_dev_alloc(ctx, 24 * k, C.byref(d)))
        try:
            base = d.value
            _lib.check(ctx, lib.gv2_memcpy_h2d(ctx, C.c_void_p(base), rows.ctypes.data_as(C.c_void_p), 8 * k))
            _lib.check(ctx, lib.gv2_memcpy_h2d(ctx, C.c_void_p(base + 8 * k), cols.ctypes.data_as(C.c_void_p), 8 * k))
            _lib.check(ctx, lib.gv2_dev_gather_cells(ctx, C.c_void_p(self.ptr), self.ld, C.c_void_p(base), C.c_void_p(base + 8 * k), k,
                                                     C.c_void_p(base + 16 * k)))
            _lib.check(ctx, lib.gv2_memcpy_d2h(ctx, out.ctypes.data_as(C.c_void_p), C.c_void_p(base + 16 * k), 8 * k))
        finally:
            lib.gv2_dev_free(ctx, d)
        return out

    def rowmax(self, row_begin: int, row_end: int, col_begin: int, col_end: int):
        m = row_end - row_begin
        mx = np.empty(m, dtype=np.float64)
        arg = np.empty(m, dtype=np.int64)
        if m == 0:
            return mx, arg
        lib, ctx = self.eng.lib, self.eng.ctx
        d = C.c_void_p()
        _lib.check(ctx, lib.gv2_dev_alloc(ctx, 16 * m, C.byref(d)))
        try:
            _lib.check(ctx, lib.gv2_dev_rowmax_block(ctx, C.c_void_p(self.ptr), self.ld, row_begin, row_end, col_begin, col_end,
                                                     C.c_void_p(d.value), C.c_void_p(d.value + 8 * m)))
            _lib.check(ctx, lib.gv2_memcpy_d2h(ctx, mx.ctypes.data_as(C.c_void_p), C.c_void_p(d.value), 8 * m))
            _lib.check(ctx, lib.gv2_memcpy_d2h(ctx, arg.ctypes.data_as(C.c_void_p), C.c_void_p(d.value + 8 * m), 8 * m))
        finally:
            lib.gv2_dev_free(ctx, d)
        return mx, arg


def _ordered_classes(TaxoGroupingList):
    seen, order = set(), []
    for t in TaxoGroupingList:
        if t not in seen:
            seen.add(t)
            order.append(t)
    return order


def intra_cells(members: np.ndarray, positions: Optional[np.ndarray] = None):
    """(rows, cols) of the class's intra list: pairs (i, j >= i) of members in row-major order
    (classification_utils.py:15-21); ``positions`` selects list positions (None = the whole list)."""
    m = len(members)
    total = m * (m + 1) // 2
    pos = np.arange(total, dtype=np.int64) if positions is None else np.asarray(positions, dtype=np.int64)
    # row a of the triangle starts at a*m - a(a-1)/2
    starts = np.arange(m, dtype=np.int64) * m - np.arange(m, dtype=np.int64) * (np.arange(m, dtype=np.int64) - 1) // 2
    a = np.searchsorted(starts, pos, side="right") - 1
    b = a + (pos - starts[a])
    return members[a], members[b]


def inter_cells(members: np.ndarray, n: int, positions: Optional[np.ndarray] = None):
    """(rows, cols) of the class's inter list: pairs (i, j > i) with exactly one of i, j in the class, row-major order
    (classification_utils.py:22-26: row i of a member lists its non-members j > i, row i of a non-member lists the
    members j > i)."""
    is_member = np.zeros(n, dtype=bool)
    is_member[members] = True
    non_members = np.flatnonzero(~is_member)
    idx = np.arange(n, dtype=np.int64)
    members_gt = len(members) - np.searchsorted(members, idx, side="right")          # members j > i
    cnt = np.where(is_member, (n - 1 - idx) - members_gt, members_gt)
    starts = np.concatenate(([0], np.cumsum(cnt)))
    total = int(starts[-1])
    pos = np.arange(total, dtype=np.int64) if positions is None else np.asarray(positions, dtype=np.int64)
    i = np.searchsorted(starts, pos, side="right") - 1
    off = pos - starts[i]
    row_is_member = is_member[i]
    j = np.empty_like(i)
    lo_nm = np.searchsorted(non_members, i, side="right")
    lo_m = np.searchsorted(members, i, side="right")
    if row_is_member.any():
        j[row_is_member] = non_members[(lo_nm + off)[row_is_member]]
    if (~row_is_member).any():
        j[~row_is_member] = members[(lo_m + off)[~row_is_member]]
    return i, j


def list_lengths(members: np.ndarray, n: int):
    m = len(members)
    return m * (m + 1) // 2, m * (n - m)


def draw_class_cells(TaxoGroupingList, N_PairwiseSimilarityScores):
    """The cells of every class's (possibly subsampled) intra and inter list, in list order, drawing from numpy's global
    legacy RNG exactly where the reference draws (classification_utils.py:34-43: intra then inter, class by class in
    first-occurrence order, only for lists longer than N).  The SVC fit that follows each class's draws in the reference
    takes its libsvm seed from the same global stream (sklearn: ``rnd.randint(np.iinfo("i").max)`` with ``random_state=None``):
    the state at that point is recorded and the draw is consumed here, so that every later draw -- the next class, the next
    bootstrap replicate -- sees the stream the reference would.
    Returns {class: (intra_rows, intra_cols, inter_rows, inter_cols, rng_state_at_fit)}."""
    taxo = np.asarray(TaxoGroupingList)
    n = len(taxo)
    out = {}
    for cls in _ordered_classes(TaxoGroupingList):
        members = np.flatnonzero(taxo == cls).astype(np.int64)
        n_intra, n_inter = list_lengths(members, n)
        p_intra = p_inter = None
        if n_intra > N_PairwiseSimilarityScores:
            p_intra = np.random.choice(n_intra, N_PairwiseSimilarityScores, replace=False)
        if n_inter > N_PairwiseSimilarityScores:
            p_inter = np.random.choice(n_inter, N_PairwiseSimilarityScores, replace=False)
        state = np.random.get_state()
        np.random.randint(np.iinfo("i").max)                       # the fit's seed draw
        out[cls] = intra_cells(members, p_intra) + inter_cells(members, n, p_inter) + (state,)
    return out


def cutoff_dict_from_cells(SimMat, cells) -> dict:
    """Lists + linear-SVC cut-off per class (classification_utils.py:45-58) from the cells of ``draw_class_cells``."""
    from sklearn.svm import SVC
    if isinstance(SimMat, DeviceMatrix):
        order = list(cells)
        rows = np.concatenate([np.concatenate((cells[c][0], cells[c][2])) for c in order]) if order else np.zeros(0, np.int64)
        cols = np.concatenate([np.concatenate((cells[c][1], cells[c][3])) for c in order]) if order else np.zeros(0, np.int64)
        flat = SimMat.gather(rows, cols)                       # one gather for all classes
    out, at = {}, 0
    for cls, (ir, ic, er, ec, fit_state) in cells.items():
        if isinstance(SimMat, DeviceMatrix):
            intra = flat[at:at + len(ir)].tolist(); at += len(ir)
            inter = flat[at:at + len(er)].tolist(); at += len(er)
        else:
            S = np.asarray(SimMat)
            intra = S[ir, ic].tolist()
            inter = S[er, ec].tolist()
        PairwiseSimilarityList = np.array(intra + inter).reshape(-1, 1)
        Labels = np.array(["Intra"] * len(intra) + ["Inter"] * len(inter))
        clf = SVC(kernel="linear", class_weight="balanced")
        if np.unique(Labels).shape[0] == 1:
            Labels[-1] = "Extra"                                # the reference's dummy second class, :53-55
        keep = np.random.get_state()
        np.random.set_state(fit_state)                             # the seed the reference's fit would have drawn
        try:
            clf.fit(PairwiseSimilarityList, Labels)
        finally:
            np.random.set_state(keep)
        out[cls] = {"PairwiseSimilarityScore_InterClass_List": inter, "PairwiseSimilarityScore_IntraClass_List": intra,
                    "CutOff": -clf.intercept_[0] / clf.coef_[0][0]}
    return out


def PairwiseSimilarityScore_Cutoff_Dict_Constructor(SimMat, TaxoGroupingList, N_PairwiseSimilarityScores) -> dict:
    """Drop-in for app/utils/classification_utils.py:10-60.  ``SimMat`` is the (N_ref, N_ref) similarity matrix as a numpy
    array or a ``DeviceMatrix``; same dict, same list order, same RNG consumption, same cut-offs."""
    return cutoff_dict_from_cells(SimMat, draw_class_cells(TaxoGroupingList, N_PairwiseSimilarityScores))


def max_similarity(SimMat, TaxoGrouping, row_begin: Optional[int] = None, n_ref: Optional[int] = None):
    """MaxSimScoreList, TaxoOfMaxSimScoreList of TaxonomicAssignmentProposerAndEvaluator (classification_utils.py:66-68).
    Host form: ``SimMat`` is the (N_ucf, N_ref) block.  Device form: ``SimMat`` is the whole DeviceMatrix, the block is
    rows [row_begin, n) x columns [0, n_ref)."""
    taxo = np.asarray(TaxoGrouping)
    if isinstance(SimMat, DeviceMatrix):
        mx, arg = SimMat.rowmax(int(row_begin), SimMat.n, 0, int(n_ref))
    else:
        S = np.asarray(SimMat)
        mx, arg = np.max(S, axis=1), np.argmax(S, axis=1)
    return mx, taxo[arg]
