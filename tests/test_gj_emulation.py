"""The generalised-Jaccard kernels (SURVEY 8a rows a1, a3, a4, a6, a7 and the fused tail of the bootstrap loops a8 / a9), run
on the CPU: tests/emul/gj_emul.cpp includes the SAME source the GPU build compiles (gravity2_b200/csrc/gj_kernels.cuh) and
runs it on host threads with the launch shapes of gv2_similarity_host / gj_fused_launch / gom_quantize_launch.  Checked
against the golden outputs of the UNMODIFIED reference (tests/golden/reference_golden.npz: all eight schemes, p = 1 and 2,
threshold 60) and, on larger ragged tables, against the oracle -- tolerance 1e-6 relative as for the GPU tests.  (cp.async
copies complete at once here and the MUFU reciprocal square root seed is an exact 1/sqrt: ring accounting and the seed's 22
bits are what the GPU tests cover.)"""
import os
import subprocess

import numpy as np
import pytest

from oracle import gravity_oracle as O
from gravity2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCHEMES = {"P": 0, "G": 1, "PG": 2, "R": 3, "RG": 4, "PR": 5, "L": 6, "PL": 7}
RTOL, ATOL = 1e-6, 1e-12


@pytest.fixture(scope="module")
def emul(emul_build):
    return emul_build("gj_emul")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


def close(a, b, rtol=RTOL, atol=ATOL):
    err = np.abs(a - b)
    assert a.shape == b.shape and (err <= atol + rtol * np.abs(b)).all(), f"worst abs {err.max():.3e}"


def similarity(emul, sig, gom, aux, scheme, power=1.0, out_kind=0, thr=0.0, apply_thr=1, ksplit=1):
    n = len(sig)
    p, g = sig.shape[1], (0 if gom is None else gom.shape[1])
    code = SCHEMES[scheme]
    need_p, need_g = code in (0, 2, 5, 7), code in (1, 2, 4)
    need_r = code in (3, 4, 5, 6, 7)
    blob = b""
    blob += np.ascontiguousarray(sig, dtype=np.float64).tobytes() if need_p else b""
    blob += np.ascontiguousarray(gom, dtype=np.float64).tobytes() if need_g else b""
    blob += np.ascontiguousarray(aux, dtype=np.float64).tobytes() if need_r else b""
    args = [n, p if need_p else 0, g if need_g else 0, code, power, out_kind, thr, apply_thr, ksplit]
    out = subprocess.run([emul, "similarity"] + [str(a) for a in args], input=blob, capture_output=True, timeout=900)
    assert out.returncode == 0, out.stderr
    return np.frombuffer(out.stdout, dtype=np.float64).reshape(n, n)


def fused(emul, sig, gom, fixed, power=1.0, out_kind=0):
    n = len(gom)
    p = 0 if sig is None else sig.shape[1]
    blob = (b"" if sig is None else np.ascontiguousarray(sig, dtype=np.float64).tobytes()) + np.ascontiguousarray(gom, dtype=np.float64).tobytes()
    out = subprocess.run([emul, "fused", str(n), str(p), str(gom.shape[1]), str(int(fixed)), str(power), str(out_kind)], input=blob,
                         capture_output=True, timeout=900)
    assert out.returncode == 0, out.stderr
    return np.frombuffer(out.stdout, dtype=np.float64).reshape(n, n)


@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("scheme,p", [("P", 1), ("G", 1), ("L", 1), ("PG", 1), ("PG", 2), ("PL", 1), ("R", 1), ("RG", 1), ("PR", 1)])
def test_emulated_similarity_kernels_equal_reference_golden(emul, golden, tag, scheme, p):
    """prepare_table_kernel + gj_min_tile_kernel + gj_epilogue_kernel == SimilarityMat_Constructor of the unmodified reference
    (schemes L / PL take the location dCor matrix as their auxiliary input, as the device path does: here the reference's own)."""
    sig, gom, spr = (golden[f"{tag}_{k}"] for k in ("sig", "gom", "spr"))
    aux = golden[f"{tag}_S_L_p1"] if scheme in ("L", "PL") else spr
    S = similarity(emul, sig, gom, aux, scheme, power=float(p))
    close(S, golden[f"{tag}_S_{scheme}_p{p}"])
    assert np.array_equal(S, S.T)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_emulated_threshold_golden(emul, golden, tag):
    S = similarity(emul, golden[f"{tag}_sig"], golden[f"{tag}_gom"], None, "PG", thr=60.0, apply_thr=1)
    close(S, golden[f"{tag}_S_PG_thr60"])


def _ragged_tables():
    t = synth.make_tables(140, 900, 6, 0.04, seed=11)
    sig = t.sig.copy()
    sig[9] = 0.0                                      # an empty genome
    sig[3] = sig[40]                                  # duplicates
    sig[17, 5] = np.nan                               # a NaN entry zeroes its genome's row (:204-205)
    rng = np.random.default_rng(5)
    gom = rng.random((140, 70)) * (rng.random((140, 70)) < 0.7)   # 70 columns: padded to 96, three stages, ragged last groups
    gom[23] = 0.0
    gom[3] = gom[40]
    return sig, gom


@pytest.mark.parametrize("scheme,power,out_kind,ksplit", [("PG", 1.0, 0, 2), ("P", 1.0, 1, 3), ("G", 2.0, 0, 1)])
def test_emulated_similarity_kernels_ragged_tables_vs_oracle(emul, scheme, power, out_kind, ksplit):
    """Two tiles per side with a ragged second one, a split K range, NaN / empty / duplicate genomes, distance output."""
    sig, gom = _ragged_tables()
    S = similarity(emul, sig, gom, None, scheme, power=power, out_kind=out_kind, ksplit=ksplit)
    with np.errstate(all="ignore"):
        ref = O.similarity_mat_fast(sig.copy(), gom, None, 0, scheme, power)
    if out_kind:
        ref = O.dist_from_sim(ref)
    close(S, ref, atol=1e-6 if out_kind else 1e-9)     # 1e-6 relative on S is 1e-6 absolute on D = 1 - S
    assert np.array_equal(S, S.T)
    assert np.all(S[9] == (1.0 if out_kind else 0.0)) if scheme != "G" else True


@pytest.mark.parametrize("fixed", [0, 1])
@pytest.mark.parametrize("with_p,power,out_kind", [(True, 1.0, 1), (True, 2.0, 0), (False, 1.0, 0)])
def test_emulated_fused_tail_vs_oracle(emul, fixed, with_p, power, out_kind):
    """gj_fused_kernel (fp32 and fixed-point variants; the latter with gom_rowmax_kernel + gom_quantize_kernel in front, as in
    the bootstrap job) against the oracle: the fixed-point image costs ~1e-8 relative on a sum, duplicates give exactly S = 1."""
    sig, gom = _ragged_tables()
    S = fused(emul, sig if with_p else None, gom, fixed, power=power, out_kind=out_kind)
    with np.errstate(all="ignore"):
        ref = O.similarity_mat_fast(sig.copy(), gom, None, 0, "PG" if with_p else "G", power)
    if out_kind:
        ref = O.dist_from_sim(ref)
    close(S, ref, rtol=2e-6 if fixed else RTOL, atol=1e-6 if out_kind else 1e-7)
    assert np.array_equal(S, S.T)
    if fixed and not with_p:      # (the PPHMM sums of this driver come from the fp32 tile kernel; the job's are exact integers)
        assert S[3, 40] == (0.0 if out_kind else 1.0)


# ------------------------------------------------------- a2: PphmmNeighbourhoodWeight != 0 (quirk Q1), reference goldens
def nbhd(emul, sig, cls, w, thr, mul=None):
    n, p = sig.shape
    blob = np.ascontiguousarray(sig, dtype=np.float64).tobytes() + np.ascontiguousarray(cls, dtype=np.uint8).tobytes()
    if mul is not None:
        blob += np.ascontiguousarray(mul, dtype=np.float64).tobytes()
    out = subprocess.run([emul, "nbhd", str(n), str(p), repr(float(w)), repr(float(thr)), str(int(mul is not None))], input=blob,
                         capture_output=True, timeout=900)
    assert out.returncode == 0, out.stderr
    return np.frombuffer(out.stdout, dtype=np.float64).reshape(n, n)


@pytest.mark.parametrize("tag", ["C", "D"])
@pytest.mark.parametrize("case", [0, 1, 2])
def test_emulated_neighbourhood_kernel_equals_reference_golden(emul, tag, case, tmp_path):
    """gj_nbhd_kernel + the ordinary combine, wired as the drop-in wires them (dropin.SimilarityMat_Constructor): the
    visit-count dependent in-place row scaling of the reference, first and second call on the same (mutated) table, against
    the outputs of the unmodified reference (reference_golden_neighbourhood.npz)."""
    from conftest import write_reference_csvs
    from gravity2_b200 import neighbourhood as NB
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_neighbourhood.npz"))
    w, thr, pw = (float(x) for x in g[f"{tag}_case{case}"])
    scheme = str(g[f"{tag}_case{case}_scheme"])
    sig = g[f"{tag}_sig"].copy()
    gom = g[f"{tag}_gom"]
    fn = write_reference_csvs(str(tmp_path), g[f"{tag}_sig"], g[f"{tag}_naive_loc"])
    cls = NB.neighbourhood_classes(*NB.read_neighbourhood_csvs(fn))
    for key in ("S", "S_second_call"):
        gjp = nbhd(emul, sig, cls, w, thr)
        S = similarity(emul, np.zeros((len(sig), 0)), gom, gjp, {"P": "R", "PG": "RG"}[scheme], power=pw)
        close(S, g[f"{tag}_case{case}_{key}"])
        NB.apply_in_place(sig, cls, w, thr, len(sig) + 1)            # what the reference leaves in the caller's table
    assert np.array_equal(sig, g[f"{tag}_case{case}_sig_after_two_calls"])


@pytest.mark.parametrize("n,p,g", [(1, 1, 1), (2, 1, 1), (129, 33, 33), (1, 40, 3)])
def test_emulated_kernels_tiny_and_odd_shapes(emul, n, p, g):
    """One genome, one column, one GOM; one row / column past a tile or stage boundary; all-zero tables."""
    rng = np.random.default_rng(n * 100 + p)
    sig = np.round(rng.gamma(2, 60, (n, p)), 1) * (rng.random((n, p)) < 0.6)
    gom = rng.random((n, g)) * (rng.random((n, g)) < 0.8)
    with np.errstate(all="ignore"):
        ref = O.similarity_mat_fast(sig.copy(), gom, None, 0, "PG", 1.0)
    close(similarity(emul, sig, gom, None, "PG", ksplit=2), ref, atol=1e-9)
    close(fused(emul, sig, gom, 1), ref, rtol=2e-6, atol=1e-7)
    assert np.all(similarity(emul, np.zeros((n, p)), None, None, "P") == 0.0)
    assert np.all(fused(emul, np.zeros((n, p)), np.zeros((n, g)), 1) == 0.0)
