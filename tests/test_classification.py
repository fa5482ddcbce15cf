"""CPU: the classification epilogue's host logic (gravity2_b200/classification.py) against the UNMODIFIED reference's
PairwiseSimilarityScore_Cutoff_Dict_Constructor (tests/golden/reference_golden_classification.npz, written by
oracle/make_golden.py classification): same class order, same lists in the same order, same cut-offs, and numpy's global
RNG left in the same state."""
import os

import numpy as np
import pytest

from gravity2_b200 import classification as CL

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_classification.npz")


@pytest.mark.parametrize("tag", ["H", "I"])
def test_cutoff_dict_equals_reference(tag):
    g = dict(np.load(GOLD, allow_pickle=False))
    S, taxo, n_pair = g[f"{tag}_S"], g[f"{tag}_taxo"], int(g[f"{tag}_npair"])
    np.random.seed(int(g[f"{tag}_seed"]))
    d = CL.PairwiseSimilarityScore_Cutoff_Dict_Constructor(S, taxo, n_pair)
    assert list(d.keys()) == [str(c) for c in g[f"{tag}_classes"]]
    for cls, v in d.items():
        assert np.array_equal(np.array(v["PairwiseSimilarityScore_IntraClass_List"]), g[f"{tag}_{cls}_intra"])
        assert np.array_equal(np.array(v["PairwiseSimilarityScore_InterClass_List"]), g[f"{tag}_{cls}_inter"])
        assert v["CutOff"] == float(g[f"{tag}_{cls}_cutoff"])
        assert isinstance(v["PairwiseSimilarityScore_IntraClass_List"], list)
    assert np.random.random() == float(g[f"{tag}_rng_next"])          # the global RNG stream stayed in step


def test_cell_lists_equal_the_literal_loop():
    """intra_cells / inter_cells against the reference's double loop (classification_utils.py:15-26) on class layouts
    with runs, singletons and an empty tail"""
    rng = np.random.default_rng(3)
    for n, k in [(1, 1), (7, 2), (40, 5), (63, 9)]:
        lab = rng.integers(0, k, n)
        for c in range(k):
            members = np.flatnonzero(lab == c)
            intra, inter = [], []
            for i in range(n):
                for j in range(i, n):
                    if lab[i] == lab[j] == c:
                        intra.append((i, j))
                    elif lab[i] != lab[j] and (lab[i] == c or lab[j] == c):
                        inter.append((i, j))
            ir, ic = CL.intra_cells(members)
            er, ec = CL.inter_cells(members, n)
            assert list(zip(ir.tolist(), ic.tolist())) == intra
            assert list(zip(er.tolist(), ec.tolist())) == inter
            assert CL.list_lengths(members, n) == (len(intra), len(inter))
            if len(inter) > 3:
                pos = rng.permutation(len(inter))[:3]
                pr, pc = CL.inter_cells(members, n, pos)
                assert [(int(a), int(b)) for a, b in zip(pr, pc)] == [inter[q] for q in pos]


def test_max_similarity_host():
    rng = np.random.default_rng(1)
    S = rng.random((6, 9))
    S[2, 3] = S[2, 7] = 2.0                           # tie: the first maximum wins, as np.argmax
    taxo = np.array([f"T{i % 4}" for i in range(9)])
    mx, tx = CL.max_similarity(S, taxo)
    assert np.array_equal(mx, S.max(axis=1)) and np.array_equal(tx, taxo[S.argmax(axis=1)]) and tx[2] == "T3"
