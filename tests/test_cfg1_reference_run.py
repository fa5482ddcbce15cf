"""BASELINE configs[0] -- the reference's own CPU-runnable case, at the shape of cli/example_run.py's run: 20 genomes, 7
taxonomic groupings, PG similarity, 100 bootstrap replicates -- against the UNMODIFIED reference.  The fixture
(tests/golden/reference_golden_cfg1.npz, oracle/make_golden.py cfg1) holds what the reference's PL1 loop
(app/src/pl1_graphs.py:78-117) produced for every replicate: the np.random.choice draw, the distance matrix and the Newick
string of DistMat2Tree(average).  CPU suite: the synthetic generator reproduces the fixture's tables, the oracle reproduces
matrices and strings of all 100 replicates bit for bit, and the kernels' own source on host threads (tests/emul) gives
matrices within the north_star tolerance and IDENTICAL DENDROGRAM TOPOLOGIES (the first 16 replicates: the emulation costs
0.7 s per replicate).  The GPU version of the same comparison, all 100 replicates, is tests/test_zz_cfg1_gpu.py."""
import os
import re
import warnings

import numpy as np
import pytest

from oracle import gravity_oracle as O
from oracle import linkage_oracle as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cfg1():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_cfg1.npz"), allow_pickle=False))


def parse_newick(s):
    """Newick string -> {frozenset(leaves below a node): branch length above it}.  Two trees have the same topology when the
    key sets agree; the root (all leaves) is not a key."""
    s = s.strip()
    assert s.endswith(";")
    pos, out = 0, {}

    def node():
        nonlocal pos
        if s[pos] == "(":
            pos += 1
            leaves = node()
            while s[pos] == ",":
                pos += 1
                leaves = leaves | node()
            assert s[pos] == ")"
            pos += 1
        else:
            m = re.match(r"[^:,()]+", s[pos:])
            leaves = frozenset([m.group(0)])
            pos += m.end()
        if s[pos] == ":":
            m = re.match(r":(-?[0-9.eE+-]+)", s[pos:])
            out[leaves] = float(m.group(1))
            pos += m.end()
        return leaves

    node()
    assert s[pos] == ";"
    return out


def same_tree(a, b, tol=4e-6):
    """Same clades, branch lengths within tol.  The strings carry 6 decimals and a branch length is a difference of two
    heights: matrices within the 1e-6 tolerance move it by up to 2e-6, the two printed roundings by another 1e-6.  (Measured
    on the emulated kernels, all 100 replicates: matrices within 4.6e-10, 98 strings identical to the last digit, the other
    two one flipped last digit.)"""
    ta, tb = parse_newick(a), parse_newick(b)
    return set(ta) == set(tb) and all(abs(ta[k] - tb[k]) <= tol for k in ta)


def test_parse_newick():
    t = parse_newick("((a:0.1,b:0.1):0.25,c:0.35);")
    assert t == {frozenset("a"): 0.1, frozenset("b"): 0.1, frozenset("ab"): 0.25, frozenset("c"): 0.35}
    assert same_tree("((a:0.1,b:0.1):0.25,c:0.35);", "(c:0.350001,(b:0.1,a:0.1):0.25);")
    assert not same_tree("((a:0.1,b:0.1):0.25,c:0.35);", "((a:0.1,c:0.1):0.25,b:0.35);")
    assert not same_tree("((a:0.1,b:0.1):0.25,c:0.35);", "((a:0.1,b:0.1):0.26,c:0.35);")


def test_cfg1_tables_are_what_the_generator_makes(cfg1):
    """bench.py, the GPU tests and the fixture see the same cfg1 tables (numpy PCG64 streams are stable across versions)."""
    from gravity2_b200 import synth
    t = synth.make_config("cfg1")
    assert np.array_equal(t.sig, cfg1["sig"]) and np.array_equal(t.loc, cfg1["loc"])
    assert list(t.taxo) == list(cfg1["taxo"]) and list(t.gom_ids) == list(cfg1["gom_ids"])
    c = synth.CONFIGS["cfg1"]
    assert cfg1["pl1_D"].shape == (c["b"], c["n"], c["n"]) and cfg1["pl1_idx"].shape == (c["b"], c["p"])
    assert len(set(cfg1["taxo"])) == c["g"] == 7
    idx = synth.reference_resample_indices(c["p"], c["b"], int(cfg1["boot_seed"]))
    assert np.array_equal(idx, cfg1["pl1_idx"])            # the reference's np.random.choice stream, draw for draw


def test_cfg1_oracle_equals_reference_for_every_replicate(cfg1):
    sig, loc, taxo, ids = cfg1["sig"], cfg1["loc"], cfg1["taxo"], [str(x) for x in cfg1["gom_ids"]]
    names = [f"v{i}" for i in range(len(sig))]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gom = O.gom_table(loc, O.gomdb(taxo, loc, ids), ids)
        assert np.array_equal(gom, cfg1["gom"])
        D0 = O.dist_from_sim(O.similarity_mat(sig.copy(), gom, loc, None, 0.0, 0, "PG", 1.0))
        assert np.array_equal(D0, cfg1["D"])
        assert O.dist_mat_to_tree(D0.copy(), names, "average", False) == str(cfg1["newick"])
        for b, idx in enumerate(cfg1["pl1_idx"]):
            D = O.pl1_bootstrap_replicate(sig, taxo, ids, idx)
            assert np.array_equal(D, cfg1["pl1_D"][b]), b
            assert O.dist_mat_to_tree(D.copy(), names, "average", False) == str(cfg1["pl1_newick"][b]), b


@pytest.fixture(scope="module")
def exes(emul_build):
    return {name: emul_build(name) for name in ("gom_emul", "bs_emul", "gj_emul", "lk_emul")}


def test_cfg1_emulated_job_and_dendrograms_equal_reference(cfg1, exes):
    """The first 16 replicates through the kernels' own source on host threads: GOM tables -> exact pair sums -> fused tail ->
    live-list UPGMA kernel -> the library's Newick printer.  Matrices within 1e-6 of the unmodified reference's, every
    dendrogram with the reference's topology and branch lengths."""
    from test_job_emulation import emulated_pl1
    from test_linkage_emulation import _native_newick, _run
    sig, taxo, ids = cfg1["sig"], cfg1["taxo"], [str(x) for x in cfg1["gom_ids"]]
    n = len(sig)
    names = [f"v{i}" for i in range(n)]
    nb = 16
    D = emulated_pl1(exes, sig, taxo, ids, cfg1["pl1_idx"][:nb])
    assert D.shape == cfg1["pl1_D"][:nb].shape
    assert np.abs(D - cfg1["pl1_D"][:nb]).max() <= 1e-6
    identical = 0
    for b in range(len(D)):
        assert np.array_equal(D[b], D[b].T) and not np.diag(D[b]).any()
        raw = _run(exes["lk_emul"], "average", D[b], 2)
        Z = raw[np.argsort(raw[:, 2], kind="mergesort")]
        L.label(Z, n)
        tree = _native_newick(Z, names)
        assert same_tree(tree, str(cfg1["pl1_newick"][b])), b
        identical += tree == str(cfg1["pl1_newick"][b])
    assert identical >= len(D) // 2                        # most strings agree to the last printed digit as well
