"""The default replicate pair-sum path of the bootstrap loops, run on the CPU:
tests/emul/bs_emul.cpp includes the SAME source the GPU build compiles (gravity2_b200/csrc/boot_sparse_kernels.cuh) and follows
boot_sparse_build / boot_sparse_build_lists / boot_sparse_launch step by step -- CSR of the 48-bit fixed-point table, the
replicate-independent pair lists (count, scan, fill), transposed multiplicities, then boot_sparse_stream_kernel (many
replicates: per-warp cp.async pipeline, byte transposes, dp4a plane sums) or boot_sparse_pairlane_kernel (<= 12 replicates:
the base matrix and the PL2 loop).  The path is exact integer arithmetic, so the check is BIT FOR BIT: Python integers restate
S[b][i][j] = sum_c w[b][c] * min(T[i][c], T[j][c]) of app/src/pl1_graphs.py:81-86 + similarity_matrix_constructor.py:167 on the
fixed-point image, and the limb recombination is one correctly rounded fused multiply-add."""
import os
import subprocess
from fractions import Fraction

import numpy as np
import pytest

from gravity2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(emul_build):
    return emul_build("bs_emul")


def ws_index(r, c):
    """gj_ws_index: element (r, c) of a 128 x 128 pair tile lives at [r/16][c/16][r%16][c%16]."""
    return ((((r >> 4) << 3) + (c >> 4)) << 8) + ((r & 15) << 4) + (c & 15)


def to_double(hi, lo, quantum):
    """bs_to_double: fma(hi, quantum * 65536, lo * quantum), each operation correctly rounded."""
    q_hi = quantum * 65536.0
    return float(Fraction(hi) * Fraction(q_hi) + Fraction(float(lo) * quantum))


@pytest.mark.parametrize("n,k,nb,density", [(1, 1, 1, 1.0), (2, 3, 2, 0.9), (140, 300, 3, 0.08), (140, 300, 10, 0.08), (130, 517, 40, 0.05), (40, 129, 1, 0.3),
                                            (150, 260, 200, 0.06)])
def test_emulated_sparse_pair_sums_are_exact(emul, n, k, nb, density):
    t = synth.make_tables(n, k, 5, density, seed=n + nb)
    tab = t.sig.copy()
    if n > 8:
        tab[7] = 0.0                                              # an empty genome
        tab[3] = tab[min(n - 1, 35)]                              # duplicates: pair sum == row sum exactly
    if tab.max() == 0.0:
        tab[0, 0] = 1.5
    rng = np.random.default_rng(nb)
    if nb == 1:
        w = np.ones((1, k), dtype=np.int32)                       # the un-resampled table as the all-ones replicate
    else:
        w = np.stack([np.bincount(rng.integers(0, k, k), minlength=k) for _ in range(nb)]).astype(np.int32)
    out = subprocess.run([emul, str(n), str(k), str(nb)], input=tab.tobytes() + w.tobytes(), capture_output=True, timeout=1800)
    assert out.returncode == 0, (out.returncode, out.stderr)
    buf = np.frombuffer(out.stdout, dtype=np.float64)
    quantum, rowsum, ws = buf[0], buf[1:1 + n], buf[1 + n:]
    nt = (n + 127) // 128
    ntl = nt * (nt + 1) // 2
    ws = ws.reshape(nb, ntl, 128 * 128)
    # the fixed-point image, as Python integers
    scale = 281474976710656.0 * (1.0 - 9.313225746154785e-10) / tab.max()
    assert quantum == 1.0 / scale
    F = np.rint(np.minimum(tab * scale, 281474976710655.0)).astype(np.int64)
    hi, lo = F >> 16, F & 0xffff
    assert all(rowsum[i] == to_double(int(hi[i].sum()), int(lo[i].sum()), quantum) for i in range(n))
    tiles = [(bi, bj) for bi in range(nt) for bj in range(bi, nt)]
    checked = 0
    pairs = [(i, j) for i in range(n) for j in range(i, n)]
    pick = rng.choice(len(pairs), size=min(len(pairs), 1500), replace=False)
    must = [(3, min(n - 1, 35)), (7, 7), (7, 9), (0, 0), (n - 1, n - 1), (0, n - 1)] if n > 9 else [(0, 0), (0, n - 1)]
    for (i, j) in [pairs[q] for q in pick] + [(min(a, b), max(a, b)) for a, b in must]:
        tix = tiles.index((i // 128, j // 128))
        cell = ws_index(i % 128, j % 128)
        mn_hi, mn_lo = np.where(F[i] <= F[j], hi[i], hi[j]), np.where(F[i] <= F[j], lo[i], lo[j])
        for b in range(nb) if nb <= 12 else rng.choice(nb, 6, replace=False):
            expect = to_double(int((w[b].astype(np.int64) * mn_hi).sum()), int((w[b].astype(np.int64) * mn_lo).sum()), quantum)
            assert ws[b, tix, cell] == expect, (i, j, b)
            checked += 1
    # duplicates: the pair sum IS the row sum of the replicate, so D = 0 exactly downstream
    if n > 8:
        d = min(n - 1, 35)
        tix, cell = tiles.index((0, 0)), ws_index(3, d)
        for b in range(min(nb, 4)):
            assert ws[b, tix, cell] == ws[b, tix, ws_index(3, 3)]
    assert checked >= min(800, len(pairs))
