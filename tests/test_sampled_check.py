"""CPU: the single-cell restatements used by the full-size GPU parity tests and by bench.py's parity_check
(oracle/sampled_check.py) agree with the literal loop-body restatements of the oracle (which are themselves pinned
to the unmodified reference's outputs by tests/test_oracle_golden.py)."""
import warnings

import numpy as np

from oracle import gravity_oracle as O
from oracle import sampled_check as SC
from gravity2_b200 import synth


def test_cells_equal_pl1_loop_body():
    t = synth.make_tables(40, 300, 4, 0.08, seed=5)
    t.sig[3] = 0.0                      # empty genome
    t.sig[39] = t.sig[2]                # duplicate genome
    np.random.seed(17)
    idx = O.bootstrap_indices(300)
    w = O.weights_from_indices(idx, 300)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b_loc = t.sig[:, idx]
        gom = O.gom_table(b_loc, O.gomdb(t.taxo, b_loc, t.gom_ids), t.gom_ids)
        D = O.pl1_bootstrap_replicate(t.sig, t.taxo, t.gom_ids, idx)
    for v in (0, 3, 2, 39, 11):
        for g, gid in enumerate(t.gom_ids):
            members = np.where(t.taxo == gid)[0]
            got = SC.gom_cell(t.sig, members, v, w)
            assert np.isclose(got, gom[v, g], rtol=1e-12, atol=1e-15) or (np.isnan(got) and np.isnan(gom[v, g]))
    rng = np.random.default_rng(1)
    for i, j in SC.sample_pairs(rng, 40, t.taxo, 30, special=[(2, 39), (3, 5), (3, 3)]):
        got = SC.pair_cell(t.sig, gom[i], gom[j], i, j, w, "PG", 1.0, True)
        assert abs(got - D[i, j]) <= 1e-12, (i, j, got, D[i, j])


def test_cells_equal_pl2_loop_body_and_base():
    t = synth.make_tables(36, 250, 4, 0.1, seed=6)
    n_ref = 28
    np.random.seed(4)
    idx = O.bootstrap_indices(250, sort=True)
    w = O.weights_from_indices(idx, 250)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S, D = O.pl2_bootstrap_replicate(t.sig, t.loc, n_ref, t.taxo[:n_ref], t.gom_ids, idx)
        b_loc = t.loc[:, idx]
        gom = O.gom_table(b_loc, O.gomdb(t.taxo[:n_ref], b_loc[:n_ref], t.gom_ids), t.gom_ids)
        gom0 = O.gom_table(t.loc, O.gomdb(t.taxo[:n_ref], t.loc[:n_ref], t.gom_ids), t.gom_ids)
        S0 = O.similarity_mat(t.sig.copy(), gom0, t.loc, None, 0.0, 0, "PG", 1.0)
    for v in (0, 30, 35):
        for g, gid in enumerate(t.gom_ids):
            members = np.where(t.taxo[:n_ref] == gid)[0]
            assert np.isclose(SC.gom_cell(t.loc, members, v, w), gom[v, g], rtol=1e-12, atol=1e-15)
            assert np.isclose(SC.gom_cell(t.loc, members, v, None), gom0[v, g], rtol=1e-12, atol=1e-15)
    for i, j in [(0, 1), (30, 2), (35, 35), (5, 29)]:
        assert abs(SC.pair_cell(t.sig, gom[i], gom[j], i, j, w, "PG", 1.0, False) - S[i, j]) <= 1e-12
        assert abs(SC.pair_cell(t.sig, gom[i], gom[j], i, j, w, "PG", 1.0, True) - D[i, j]) <= 1e-12
        assert abs(SC.pair_cell(t.sig, gom0[i], gom0[j], i, j, None, "PG", 1.0, False) - S0[i, j]) <= 1e-12


def test_pair_cell_negative_and_nan_entries():
    """threshold 0 zeroes negative scores (similarity_matrix_constructor.py:139-140); a NaN zeroes the row only where
    its column was drawn"""
    sig = np.array([[1.0, -2.0, 3.0, 0.0], [2.0, 5.0, 1.0, 4.0], [np.nan, 1.0, 1.0, 1.0]])
    ref = O.similarity_mat(sig.copy(), None, None, None, 0.0, 0, "P", 1.0)
    for i in range(3):
        for j in range(3):
            assert abs(SC.pair_cell(sig, None, None, i, j, None, "P", 1.0, False) - ref[i, j]) <= 1e-15
    w = np.array([0, 2, 1, 1])
    ref_w = O.similarity_mat(sig[:, np.repeat(np.arange(4), w)].copy(), None, None, None, 0.0, 0, "P", 1.0)
    for i in range(3):
        for j in range(3):
            assert abs(SC.pair_cell(sig, None, None, i, j, w, "P", 1.0, False) - ref_w[i, j]) <= 1e-15
