"""BASELINE configs[0] on the GPU: the run of cli/example_run.py's shape (20 genomes, 7 taxonomic groupings, PG, 100 bootstrap
replicates) through the drop-in functions and the C ABI, against what the UNMODIFIED reference produced for the same tables
and the same np.random.choice stream (tests/golden/reference_golden_cfg1.npz, oracle/make_golden.py cfg1): GOM table rel
1e-9, every distance matrix within 1e-6 (north_star), every dendrogram with the reference's topology.  The CPU suite runs the
same comparison on the oracle and on the emulated kernels (tests/test_cfg1_reference_run.py, which also holds the Newick
comparison helpers).  The file sorts after the rest of the GPU suite on purpose."""
import os

import numpy as np
import pytest

from test_cfg1_reference_run import same_tree

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAYLOAD = {"N_Bootstrap": 100, "SimilarityMeasurementScheme": "PG", "p": 1.0, "Dendrogram_LinkageMethod": "average",
           "PphmmNeighbourhoodWeight": 0.0, "PphmmSigScoreThreshold": 0, "N_CPUs": 1}


@pytest.fixture(scope="module")
def cfg1():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_cfg1.npz"), allow_pickle=False))


def test_cfg1_base_matrix_gom_table_and_tree(cfg1):
    """The un-resampled tables through the reference's own call sequence (pl1_graphs.py:40-54, 91-100, 110-117)."""
    import gravity2_b200 as G
    sig, loc, taxo, ids = cfg1["sig"], cfg1["loc"], cfg1["taxo"], [str(x) for x in cfg1["gom_ids"]]
    names = [f"v{i}" for i in range(len(sig))]
    gom = G.GOMSignatureTable_Constructor(loc, G.GOMDB_Constructor(taxo, loc, ids), ids, PAYLOAD)
    assert gom.shape == cfg1["gom"].shape and gom.dtype == np.float64
    assert np.all(np.abs(gom - cfg1["gom"]) <= 1e-12 + 1e-9 * np.abs(cfg1["gom"]))
    sig_in = sig.copy()
    D = G.dist_mat_constructor(sig_in, gom, loc, None, PAYLOAD)
    assert np.array_equal(sig_in, sig)                        # weight 0, threshold 0: the caller's table is left as it was
    assert D.shape == cfg1["D"].shape and D.dtype == np.float64
    assert np.abs(D - cfg1["D"]).max() <= 1e-6 and np.array_equal(D, D.T)
    assert same_tree(G.DistMat2Tree(D.copy(), names, "average", False), str(cfg1["newick"]))


def test_cfg1_all_100_replicates_matrices_and_dendrograms(cfg1):
    """The batched PL1 entry points with the reference's RNG stream: matrices, then (same seed again) the device tree path."""
    import gravity2_b200 as G
    sig, taxo, ids = cfg1["sig"], cfg1["taxo"], [str(x) for x in cfg1["gom_ids"]]
    names = [f"v{i}" for i in range(len(sig))]
    nb = len(cfg1["pl1_D"])
    np.random.seed(int(cfg1["boot_seed"]))
    mats = [m.copy() for m in G.bootstrap_distance_matrices(sig, None, taxo, ids, PAYLOAD, pipeline="pl1")]
    after = np.random.randint(0, 1 << 30)
    assert len(mats) == nb == 100
    for b in range(nb):
        assert np.abs(mats[b] - cfg1["pl1_D"][b]).max() <= 1e-6, b
        assert np.array_equal(mats[b], mats[b].T) and not np.diag(mats[b]).any()
    np.random.seed(int(cfg1["boot_seed"]))
    trees = list(G.bootstrap_newick(sig, None, taxo, ids, PAYLOAD, names, pipeline="pl1"))
    assert after == np.random.randint(0, 1 << 30)             # both consumed the stream exactly as the reference's loop did
    assert len(trees) == nb
    for b in range(nb):
        assert same_tree(trees[b], str(cfg1["pl1_newick"][b])), b


def test_cfg1_unmodified_loop_body_on_the_drop_ins(cfg1):
    """app/src/pl1_graphs.py:78-117 as written, one replicate per iteration, its callees swapped for the drop-ins."""
    import gravity2_b200 as G
    sig, taxo, ids = cfg1["sig"], cfg1["taxo"], [str(x) for x in cfg1["gom_ids"]]
    n, p = sig.shape
    names = [f"v{i}" for i in range(n)]
    np.random.seed(int(cfg1["boot_seed"]))
    for b in range(6):
        idx = np.random.choice(list(range(p)), p, replace=True)
        assert np.array_equal(idx, cfg1["pl1_idx"][b])
        b_sig = sig[:, idx]
        b_loc = sig[:, idx]
        bdb = G.GOMDB_Constructor(taxo, b_loc, ids)
        bgom = G.GOMSignatureTable_Constructor(b_loc, bdb, ids, PAYLOAD, b)
        D = G.dist_mat_constructor(b_sig, bgom, b_loc, None, PAYLOAD)
        assert np.abs(D - cfg1["pl1_D"][b]).max() <= 1e-6, b
        assert same_tree(G.DistMat2Tree(D, names, "average", False), str(cfg1["pl1_newick"][b])), b
