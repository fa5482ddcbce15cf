"""The tensor-core replicate pair sums (dense tables), run on the CPU: tests/emul/mma_emul.cpp includes
the SAME source the GPU build compiles (gravity2_b200/csrc/boot_mma_kernels.cuh: 16 producer warps, a TMA warp, an MMA-issuing
warp, mbarrier hand-overs, SWIZZLE_128B operand tiles, TMEM accumulators) and supplies a functional model of mbarrier / TMA /
tcgen05 behind the kernel's PTX wrappers, written from the PTX ISA's description (the descriptors are decoded, not assumed).
The path is exact integer arithmetic, so the check is BIT FOR BIT against Python integers:
    S[b][i][j] = sum_c w[b][c] * min(T[i][c], T[j][c])      (app/src/pl1_graphs.py:81-86 + similarity_matrix_constructor.py:167)
on the 32-bit fixed-point image, low and high half-word passes recombined as the kernel's epilogue does.  Two mutants show the
check has teeth: a cp.async wait that is one group too weak, and a tensor core that ignores the swizzle mode.  The compiled
kernel is checked on the B200 (test_gpu_parity.py::test_tensor_core_path_is_exact_integer_arithmetic)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from gravity2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(emul_build):
    return emul_build("mma_emul")


def ws_index(r, c):
    return ((((r >> 4) << 3) + (c >> 4)) << 8) + ((r & 15) << 4) + (c & 15)


def case(n, k, nb, density):
    tab = synth.make_tables(n, k, 5, density, seed=n + nb).sig.copy()
    if n > 8:
        tab[7] = 0.0                                              # an empty genome
        tab[3] = tab[min(n - 1, 35)]                              # duplicates: pair sum == row sum exactly
    if tab.max() == 0.0:
        tab[0, 0] = 1.5
    rng = np.random.default_rng(nb)
    w = np.stack([np.bincount(rng.integers(0, k, k), minlength=k) for _ in range(nb)]).astype(np.int32)
    return tab, w


def mismatches(exe, tab, w, properties=True):
    n, k = tab.shape
    nb = len(w)
    out = subprocess.run([exe, str(n), str(k), str(nb)], input=tab.tobytes() + w.tobytes(), capture_output=True, timeout=1800)
    assert out.returncode == 0, (out.returncode, out.stderr[-500:])
    buf = np.frombuffer(out.stdout, dtype=np.float64)
    quantum, overflow = buf[0], buf[1]
    rows, ws = buf[2:2 + nb * n].reshape(nb, n), buf[2 + nb * n:]
    nt = (n + 127) // 128
    tiles = [(a, b) for a in range(nt) for b in range(a, nt)]
    ws = ws.reshape(nb, len(tiles), 128 * 128)
    assert overflow == 0 and quantum == tab.max() / 4294967295.0
    F = np.rint(np.minimum(tab * (4294967295.0 / tab.max()), 4294967295.0)).astype(np.int64)    # table_to_fixed_kernel
    W = w.astype(np.int64)
    bad = 0
    for i in range(n):
        mn = np.minimum(F[i][None, :], F[i:])                                    # (n - i, k) exact integers
        lo, hi = (mn & 0xffff) @ W.T, (mn >> 16) @ W.T                           # (n - i, nb): the two passes' plane sums
        want = lo.astype(np.float64) * quantum + hi.astype(np.float64) * 65536.0 * quantum      # store, then += as the epilogue
        got = np.stack([ws[:, tiles.index((i // 128, j // 128)), ws_index(i % 128, j % 128)] for j in range(i, n)])
        bad += int((got != want).sum())
        assert np.array_equal(rows[:, i], got[0])                                # weighted row sums = the diagonal pairs
    if n > 35 and properties:
        assert np.array_equal(ws[:, tiles.index((0, 0)), ws_index(3, 35)], rows[:, 3])          # duplicate genomes
        assert not rows[:, 7].any()
    return bad


@pytest.mark.parametrize("n,k,nb,density", [(1, 1, 1, 1.0), (40, 200, 20, 0.3), (130, 200, 20, 0.2), (40, 700, 17, 0.1), (20, 100, 270, 0.5)])
def test_emulated_tensor_core_pair_sums_are_exact(emul, n, k, nb, density):
    """one genome / column / replicate; two pipeline stages; three genome tiles; six column chunks (the cp.async ring and both
    mbarrier phases wrap); two replicate blocks of 144 with padded multiplicity rows."""
    assert mismatches(emul, *case(n, k, nb, density)) == 0


@pytest.mark.parametrize("flag", ["MMA_EMUL_WEAK_WAIT", "MMA_EMUL_LINEAR_OPERANDS"])
def test_mutants_are_caught(tmp_path, flag):
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    exe = str(tmp_path / "mutant")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-ffp-contract=off", "-D" + flag, "-o", exe,
                    os.path.join(ROOT, "tests", "emul", "mma_emul.cpp")], check=True)
    assert mismatches(exe, *case(40, 700, 17, 0.1), properties=False) > 0


def test_emulated_tensor_core_extreme_digits_and_multiplicities(emul):
    """Every digit plane at 255 (all entries at the table's maximum -> fixed point 2^32 - 1), multiplicities at the uint8 limit,
    a zero replicate: the int32 TMEM accumulators and the fp64 recombination stay exact."""
    n, k = 9, 130
    tab = np.full((n, k), 7.5)
    tab[4, 100:] = 0.0
    tab[5, :] = 7.5 / 3.0
    w = np.ones((3, k), dtype=np.int32)
    w[0, ::2] = 255
    w[1, :] = 0
    w[2, 129] = 255
    assert mismatches(emul, tab, w, properties=False) == 0


def test_emulated_weight_overflow_is_flagged(emul):
    """A multiplicity above 255 (or negative) cannot enter the uint8 operand: weights_to_u8_kernel raises the flag the job reads
    before any delivery (api.cu validates the multiplicities up front; this is the kernel-side guard)."""
    tab = np.full((2, 5), 1.0)
    for bad in (256, -1):
        w = np.ones((1, 5), dtype=np.int32)
        w[0, 2] = bad
        out = subprocess.run([emul, "2", "5", "1"], input=tab.tobytes() + w.tobytes(), capture_output=True, timeout=600)
        assert out.returncode == 0 and np.frombuffer(out.stdout, dtype=np.float64)[1] == 1.0
