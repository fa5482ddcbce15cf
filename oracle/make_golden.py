"""Generate tests/golden/*.npz by running the UNMODIFIED reference (Mayne941/gravity2)
imported from /root/reference.  Runs only in the build container (the GPU box has no
/root/reference); the fixtures it writes are committed and are what pins the oracle.

    python oracle/make_golden.py

Cosmetic-only imports the container lacks (termcolor, alive_progress: oracle/stubs/;
plotly, Bio, matplotlib: empty stand-ins created below) are stubbed; none of them is on
the arithmetic path.
"""
import os
import sys
import tempfile
import types
import warnings
from unittest import mock

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("GRAVITY_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "stubs"), REF, REPO]

for name in ["plotly", "plotly.express", "Bio", "Bio.Phylo", "matplotlib", "matplotlib.pyplot",
             "matplotlib.colors", "matplotlib.patches", "matplotlib.lines", "ete3"]:
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = mock.MagicMock()

from app.utils.similarity_matrix_constructor import SimilarityMat_Constructor  # noqa: E402
from app.utils.parallel_gom_sig_generator import generate_gom_sigs  # noqa: E402
from app.utils.gomdb_constructor import GOMDB_Constructor  # noqa: E402
from app.utils.dcor import dcor, cent_dist  # noqa: E402
from app.utils.dist_mat_to_tree import DistMat2Tree  # noqa: E402
from app.utils.shared_pphmm_graphs import shared_pphmm_ratio, shared_norm_pphmm_ratio  # noqa: E402

from gravity2_b200.synth import make_tables  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def write_csvs(tmp, sig, loc):
    """The two CSVs SimilarityMat_Constructor insists on, laid out as
    app/src/pl1_graphs.py:307-319 writes them."""
    n, p = sig.shape
    names = [f"v{i}" for i in range(n)]
    cols = [f"PPHMM {c}" for c in range(p)]
    df = pd.DataFrame(sig, columns=cols)
    df.insert(0, "Virus name", names)
    f_sig = os.path.join(tmp, "sigs.csv")
    df.to_csv(f_sig)
    dl = pd.DataFrame(np.abs(loc), columns=cols)
    dl.insert(0, "Virus name", names)
    f_loc = os.path.join(tmp, "locs.csv")
    dl.to_csv(f_loc, index=False)
    return {"PphmmLocs": f_loc, "PphmmAndGomSigs": f_sig}


def ref_gom_table(loc, db, ids):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.column_stack([generate_gom_sigs(g, loc, db) for g in ids])


def main():
    gold = {}
    # --- known answers the reference itself holds
    gold["dcor_kat"] = np.float64(dcor(np.matrix("1;2;3;4;5"), np.matrix("1;2;9;4;4")))
    gold["cent_dist_123456"] = cent_dist(np.array([[1, 2], [3, 4], [5, 6]]))
    fx_loc = np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1]])
    fx_db = {"GOM1": np.array([[1, 0, 1], [0, 1, 0], [1, 1, 1]]),
             "GOM2": np.array([[0, 1, 0], [1, 0, 1], [0, 0, 0]])}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gold["gomfix_GOM1"] = np.array(generate_gom_sigs("GOM1", fx_loc, fx_db))
        gold["gomfix_GOM2"] = np.array(generate_gom_sigs("GOM2", fx_loc, fx_db))
    D3 = np.array([[0.0, 0.2, 0.4], [0.2, 0.0, 0.6], [0.4, 0.6, 0.0]])
    gold["newick_3x3_single"] = np.array(DistMat2Tree(D3.copy(), ["A", "B", "C"], "single", False))

    with tempfile.TemporaryDirectory() as tmp:
        # --- small synthetic case A (sparse, PL1-like) and B (denser, with empty genome + duplicate)
        for tag, (n, p, g, dens, seed) in {"A": (14, 60, 4, 0.25, 101), "B": (10, 33, 3, 0.6, 202)}.items():
            t = make_tables(n, p, g, dens, seed)
            sig, loc = t.sig, t.loc
            if tag == "B":
                sig[3] = 0.0; loc[3] = 0.0            # all-zero genome (Q5: S[i,i] = 0)
                sig[7] = sig[2]; loc[7] = loc[2]      # duplicate genome (S = 1 off-diagonal)
            fn = write_csvs(tmp, sig, loc)
            db = GOMDB_Constructor(t.taxo, loc, t.gom_ids)
            gom = ref_gom_table(loc, db, t.gom_ids)
            gold[f"{tag}_sig"], gold[f"{tag}_loc"], gold[f"{tag}_gom"] = sig, loc, gom
            gold[f"{tag}_taxo"] = t.taxo
            gold[f"{tag}_gom_ids"] = np.array(t.gom_ids)
            spr = np.random.default_rng(seed).random((n, n))
            spr = (spr + spr.T) / 2
            gold[f"{tag}_spr"] = spr
            for scheme in ["P", "G", "L", "PG", "PL", "R", "RG", "PR"]:
                for pw in ([1.0, 2.0] if scheme == "PG" else [1.0]):
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        S = SimilarityMat_Constructor(sig.copy(), gom, loc, spr.copy(), 0.0, 0, scheme, pw, fn)
                    gold[f"{tag}_S_{scheme}_p{int(pw)}"] = np.array(S)
            # threshold > 0 (weight == 0): idempotent in-place zeroing (Q1)
            sig_thr = sig.copy()
            S = SimilarityMat_Constructor(sig_thr, gom, loc, None, 0.0, 60, "PG", 1.0, fn)
            gold[f"{tag}_S_PG_thr60"] = np.array(S)
            gold[f"{tag}_sig_after_thr60"] = sig_thr
            # presence/absence matrices (a14, a15) through the reference's CSV reader
            order = list(range(n))
            keep = [i for i in order if np.count_nonzero(sig[i])]     # a15 raises on empty genomes
            gold[f"{tag}_shared"] = np.array(shared_pphmm_ratio(order, fn, None), dtype=np.int64)
            gold[f"{tag}_shared_norm_keep"] = np.array(keep)
            gold[f"{tag}_shared_norm"] = shared_norm_pphmm_ratio(keep, fn, None)
            # --- bootstrap loop bodies with the reference's own RNG call (Q4)
            np.random.seed(1234 + seed)
            idx_pl1, D_pl1, idx_pl2, S_pl2 = [], [], [], []
            for _ in range(3):
                # PL1, app/src/pl1_graphs.py:81-107 (signature table doubles as location table, Q3)
                idx = np.random.choice(list(range(p)), p, replace=True)
                b_sig = sig[:, idx]; b_loc = sig[:, idx]
                bdb = GOMDB_Constructor(t.taxo, b_loc, t.gom_ids)
                bgom = ref_gom_table(b_loc, bdb, t.gom_ids)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    S = SimilarityMat_Constructor(b_sig, bgom, b_loc, None, 0.0, 0, "PG", 1.0, fn)
                D = 1 - S; D[D < 0] = 0
                idx_pl1.append(idx); D_pl1.append(D)
            n_ref = n - 3
            for _ in range(3):
                # PL2, app/src/virus_classification.py:295-315 (sorted idx, real location table)
                idx = sorted(np.random.choice(list(range(p)), p, replace=True))
                b_sig = sig[:, idx]; b_loc = loc[:, idx]
                bdb = GOMDB_Constructor(t.taxo[:n_ref], b_loc[:n_ref], t.gom_ids)
                bgom = ref_gom_table(b_loc, bdb, t.gom_ids)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    S = SimilarityMat_Constructor(b_sig, bgom, b_loc, None, 0.0, 0, "PG", 1.0, fn)
                idx_pl2.append(np.array(idx)); S_pl2.append(np.array(S))
            gold[f"{tag}_boot_seed"] = np.int64(1234 + seed)
            gold[f"{tag}_pl1_idx"] = np.array(idx_pl1)
            gold[f"{tag}_pl1_D"] = np.array(D_pl1)
            gold[f"{tag}_pl2_nref"] = np.int64(n_ref)
            gold[f"{tag}_pl2_idx"] = np.array(idx_pl2)
            gold[f"{tag}_pl2_S"] = np.array(S_pl2)
            D0 = 1 - gold[f"{tag}_S_PG_p1"]; D0[D0 < 0] = 0
            gold[f"{tag}_newick_avg"] = np.array(DistMat2Tree(D0.copy(), [f"v{i}" for i in range(n)], "average", False))
    np.savez_compressed(os.path.join(OUT, "reference_golden.npz"), **gold)
    print("wrote", os.path.join(OUT, "reference_golden.npz"), len(gold), "arrays")


def main_neighbourhood():
    """PphmmNeighbourhoodWeight != 0: the unmodified reference on
    tables with runs of >= 4 consecutive hits in location order, one of them reaching the last column."""
    gold = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag, (n, p, g, dens, seed) in {"C": (12, 48, 3, 0.55, 303), "D": (9, 40, 3, 0.35, 404)}.items():
            t = make_tables(n, p, g, dens, seed)
            sig, loc = t.sig, t.loc
            # naive location table of the CSV: |loc|, increasing with the column number so that the location
            # order is (mostly) the column order and hit runs survive the sort; the last column sorts last
            rng = np.random.default_rng(seed)
            naive = np.where(sig != 0, (np.arange(p)[None, :] * 100.0 + rng.integers(1, 90, (n, p))), 0.0)
            sig[1, -6:] = [55.5, 80.1, 20.2, 90.3, 10.4, 70.7]       # run that reaches the last column
            naive[1, -6:] = (np.arange(p)[-6:] * 100.0 + 7)
            sig[2, :5] = [30.0, 45.0, 60.0, 75.0, 15.0]              # run starting at the (skipped) first column
            naive[2, :5] = (np.arange(5) * 100.0 + 3)
            sig[2, 5] = 0.0; naive[2, 5] = 0.0
            if tag == "D":
                sig[4] = 0.0; naive[4] = 0.0                         # all-zero genome
            fn = write_csvs(tmp, sig, naive)
            db = GOMDB_Constructor(t.taxo, loc, t.gom_ids)
            gom = ref_gom_table(loc, db, t.gom_ids)
            gold[f"{tag}_sig"], gold[f"{tag}_naive_loc"], gold[f"{tag}_gom"] = sig.copy(), naive, gom
            for k, (w, thr, scheme, pw) in enumerate([(0.05, 0, "P", 1.0), (0.05, 30, "PG", 1.0), (0.3, 12.5, "PG", 2.0)]):
                s_in = sig.copy()
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    S = SimilarityMat_Constructor(s_in, gom, loc, None, w, thr, scheme, pw, fn)
                    S2 = SimilarityMat_Constructor(s_in, gom, loc, None, w, thr, scheme, pw, fn)   # second call starts from the mutated table
                gold[f"{tag}_case{k}"] = np.array([w, thr, pw])
                gold[f"{tag}_case{k}_scheme"] = np.array(scheme)
                gold[f"{tag}_case{k}_S"] = np.array(S)
                gold[f"{tag}_case{k}_S_second_call"] = np.array(S2)
                gold[f"{tag}_case{k}_sig_after_two_calls"] = s_in
    np.savez_compressed(os.path.join(OUT, "reference_golden_neighbourhood.npz"), **gold)
    print("wrote reference_golden_neighbourhood.npz", len(gold), "arrays")


class _AutoMock:
    """meta-path finder: any module under these prefixes that the container lacks becomes a MagicMock (none is on the
    arithmetic path: sequence IO, plotting, shell wrappers)"""
    PREFIXES = ("Bio", "ete3", "plotly", "matplotlib", "dendropy", "seaborn")

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        if name.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(name, self)
        return None

    def create_module(self, spec):
        m = mock.MagicMock()
        m.__path__ = []
        m.__name__ = spec.name
        return m

    def exec_module(self, module):
        pass


def main_sort_pphmm():
    """a16: the PPHMM x PPHMM binary Jaccard of RefVirusAnnotator.sort_pphmm (app/src/ref_virus_annotator.py:143-158).
    The UNMODIFIED method runs on a stand-in `self` holding only the signature table; the DistMat2Tree call that follows
    the similarity loop (:162) is intercepted to capture `1 - SimMat_traits`, then the method is left (the rest rebuilds
    HH-suite databases with shell commands)."""
    for k in [k for k in sys.modules if k.split(".")[0] in _AutoMock.PREFIXES]:
        del sys.modules[k]
    sys.meta_path.insert(0, _AutoMock())
    import app.src.ref_virus_annotator as rva

    class _Captured(Exception):
        pass

    gold = {}
    for tag, (n, p, dens, seed, empty_col) in {"E": (23, 40, 0.25, 505, False), "F": (17, 33, 0.15, 606, True)}.items():
        t = make_tables(n, p, 3, dens, seed)
        sig = t.sig
        if empty_col:
            sig[:, 5] = 0.0                      # two empty PPHMM columns: 0/0 -> NaN (the reference has no zero guard)
            sig[:, 20] = 0.0
        else:
            for c in np.where((sig != 0).sum(0) == 0)[0]:
                sig[c % n, c] = 12.5             # no empty column: linkage needs finite distances
        box = {}

        def fake_tree(DistMat, LeafList, Dendrogram_LinkageMethod, do_logscale):
            box["D"] = np.array(DistMat)
            box["leaves"] = list(LeafList)
            raise _Captured()

        me = types.SimpleNamespace(PPHMMSignatureTable=sig.copy(), payload={"LogscaleDendrogram": False})
        with mock.patch.object(rva, "DistMat2Tree", fake_tree), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                rva.RefVirusAnnotator.sort_pphmm(me)
            except _Captured:
                pass
        gold[f"{tag}_sig"] = sig
        gold[f"{tag}_traits_D"] = box["D"]                       # 1 - SimMat_traits, (P, P), NaN for empty-vs-anything
        assert box["leaves"] == list(range(p))
        if not empty_col:
            gold[f"{tag}_traits_newick"] = np.array(DistMat2Tree(box["D"].copy(), box["leaves"], "average", False))
    np.savez_compressed(os.path.join(OUT, "reference_golden_sort_pphmm.npz"), **gold)
    print("wrote reference_golden_sort_pphmm.npz", len(gold), "arrays")


def main_classification():
    """next-2: PairwiseSimilarityScore_Cutoff_Dict_Constructor (app/utils/classification_utils.py:10-60), UNMODIFIED, on
    seeded similarity matrices: lists longer and shorter than N_PairwiseSimilarityScores, a single-member class, and the
    state of numpy's global RNG afterwards (the drop-in must consume it identically)."""
    for k in [k for k in sys.modules if k.split(".")[0] in _AutoMock.PREFIXES]:
        del sys.modules[k]
    sys.meta_path.insert(0, _AutoMock())
    from app.utils.classification_utils import PairwiseSimilarityScore_Cutoff_Dict_Constructor as ref_fn
    gold = {}
    for tag, (n, classes, n_pair, seed) in {"H": (60, [25, 20, 14, 1], 40, 11), "I": (33, [12, 11, 10], 1000, 12)}.items():
        rng = np.random.default_rng(seed)
        taxo = np.array(sum([[f"C{k}"] * m for k, m in enumerate(classes)], []))
        taxo = taxo[rng.permutation(n)]
        lab = np.array([int(t[1:]) for t in taxo])
        S = rng.random((n, n)) * 0.3 + 0.55 * (lab[:, None] == lab[None, :])
        S = (S + S.T) / 2
        np.fill_diagonal(S, 1.0)
        np.random.seed(100 + seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d = ref_fn(S, taxo, n_pair)
        gold[f"{tag}_S"], gold[f"{tag}_taxo"], gold[f"{tag}_npair"], gold[f"{tag}_seed"] = S, taxo, np.int64(n_pair), np.int64(100 + seed)
        gold[f"{tag}_classes"] = np.array(list(d.keys()))
        for cls, v in d.items():
            gold[f"{tag}_{cls}_intra"] = np.array(v["PairwiseSimilarityScore_IntraClass_List"], dtype=np.float64)
            gold[f"{tag}_{cls}_inter"] = np.array(v["PairwiseSimilarityScore_InterClass_List"], dtype=np.float64)
            gold[f"{tag}_{cls}_cutoff"] = np.float64(v["CutOff"])
        gold[f"{tag}_rng_next"] = np.float64(np.random.random())
    np.savez_compressed(os.path.join(OUT, "reference_golden_classification.npz"), **gold)
    print("wrote reference_golden_classification.npz", len(gold), "arrays")


def main_grouping():
    """next-3: VirusGrouping_Estimator (app/utils/virus_grouping_estimator.py:74-89), UNMODIFIED, on block-structured distance
    matrices."""
    from app.utils.virus_grouping_estimator import VirusGrouping_Estimator as ref_fn
    gold = {}
    for tag, (n, k, seed) in {"J": (48, 4, 21), "K": (30, 3, 22)}.items():
        rng = np.random.default_rng(seed)
        lab = rng.integers(0, k, n)
        taxo = np.array([f"G{x}" for x in lab])
        D = rng.random((n, n)) * 0.25 + 0.6 * (lab[:, None] != lab[None, :])
        D = (D + D.T) / 2
        np.fill_diagonal(D, 0.0)
        groups, cut, score, u1, u2 = ref_fn(D, "average", taxo)
        gold[f"{tag}_D"], gold[f"{tag}_taxo"] = D, taxo
        gold[f"{tag}_groups"], gold[f"{tag}_stats"] = np.asarray(groups), np.array([cut, score, u1, u2])
    np.savez_compressed(os.path.join(OUT, "reference_golden_grouping.npz"), **gold)
    print("wrote reference_golden_grouping.npz", len(gold), "arrays")


def main_linkage():
    """next-1, every linkage method api_classes.py:123 allows: the UNMODIFIED DistMat2Tree (app/utils/dist_mat_to_tree.py:5-21)
    on the distance matrices of the A / B fixtures (1 - S_PG, as dist_mat_constructor builds them) and on a block-structured
    matrix with exact ties; centroid and median give non-monotone heights (negative branch lengths in the Newick string)."""
    base = np.load(os.path.join(OUT, "reference_golden.npz"))
    gold = {}
    mats = {}
    for tag in ("A", "B"):
        D = 1 - base[f"{tag}_S_PG_p1"]
        D[D < 0] = 0
        mats[tag] = D
    rng = np.random.default_rng(77)
    grp = rng.integers(0, 5, 40)
    T = np.where(grp[:, None] == grp[None, :], np.round(rng.random((40, 40)) * 0.5, 2), 1.0)
    T = np.triu(T, 1)
    mats["T"] = T + T.T
    for tag, D in mats.items():
        gold[f"{tag}_D"] = D
        for method in ("single", "complete", "average", "weighted", "centroid", "median", "ward"):
            for logscale in (False, True):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    nwk = DistMat2Tree(D.copy(), [f"v{i}" for i in range(len(D))], method, logscale)
                gold[f"{tag}_{method}_{int(logscale)}"] = np.array(nwk)
    np.savez_compressed(os.path.join(OUT, "reference_golden_linkage.npz"), **gold)
    print("wrote reference_golden_linkage.npz", len(gold), "arrays")


def main_cfg1():
    """BASELINE configs[0] at the shape of cli/example_run.py's run (20 genomes, 7 taxonomic groupings, 100 bootstrap
    replicates; synthetic tables of gravity2_b200.synth cfg1 -- the example's hmmscan stage cannot run here): the base PG
    matrix, then the UNMODIFIED PL1 loop body (app/src/pl1_graphs.py:78-117) for every replicate -- np.random.choice draw,
    sig[:, idx] as signature AND location table (Q3), GOM database and table rebuilt, SimilarityMat_Constructor, 1 - S clamped,
    DistMat2Tree(average) -- keeping every replicate's distance matrix and Newick string."""
    from gravity2_b200.synth import CONFIGS, SEED_BASE
    c = CONFIGS["cfg1"]
    t = make_tables(c["n"], c["p"], c["g"], c["density"], SEED_BASE + 1)
    sig, loc, n, p = t.sig, t.loc, c["n"], c["p"]
    names = [f"v{i}" for i in range(n)]
    gold = {"sig": sig, "loc": loc, "taxo": t.taxo, "gom_ids": np.array(t.gom_ids), "boot_seed": np.int64(SEED_BASE + 1)}
    with tempfile.TemporaryDirectory() as tmp:
        fn = write_csvs(tmp, sig, loc)
        db = GOMDB_Constructor(t.taxo, loc, t.gom_ids)
        gom = ref_gom_table(loc, db, t.gom_ids)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            S = SimilarityMat_Constructor(sig.copy(), gom, loc, None, 0.0, 0, "PG", 1.0, fn)
        D0 = 1 - np.array(S); D0[D0 < 0] = 0
        gold["gom"], gold["D"] = gom, D0
        gold["newick"] = np.array(DistMat2Tree(D0.copy(), names, "average", False))
        np.random.seed(int(gold["boot_seed"]))
        idxs, Ds, nwk = [], [], []
        for _ in range(c["b"]):
            idx = np.random.choice(list(range(p)), p, replace=True)
            b_sig = sig[:, idx]; b_loc = sig[:, idx]
            bdb = GOMDB_Constructor(t.taxo, b_loc, t.gom_ids)
            bgom = ref_gom_table(b_loc, bdb, t.gom_ids)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                S = SimilarityMat_Constructor(b_sig, bgom, b_loc, None, 0.0, 0, "PG", 1.0, fn)
            D = 1 - S; D[D < 0] = 0
            idxs.append(idx); Ds.append(np.array(D))
            nwk.append(DistMat2Tree(D.copy(), names, "average", False))
    gold["pl1_idx"], gold["pl1_D"], gold["pl1_newick"] = np.array(idxs), np.array(Ds), np.array(nwk)
    np.savez_compressed(os.path.join(OUT, "reference_golden_cfg1.npz"), **gold)
    print("wrote reference_golden_cfg1.npz", len(gold), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cfg1":
        main_cfg1()
    elif len(sys.argv) > 1 and sys.argv[1] == "linkage":
        main_linkage()
    elif len(sys.argv) > 1 and sys.argv[1] == "grouping":
        main_grouping()
    elif len(sys.argv) > 1 and sys.argv[1] == "classification":
        main_classification()
    elif len(sys.argv) > 1 and sys.argv[1] == "neighbourhood":
        main_neighbourhood()
    elif len(sys.argv) > 1 and sys.argv[1] == "sort_pphmm":
        main_sort_pphmm()
    else:
        main()
