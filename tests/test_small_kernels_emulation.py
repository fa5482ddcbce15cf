"""The classification-epilogue kernels and the device-RNG kernel, run on the CPU: tests/emul/small_emul.cpp
includes the SAME source the GPU build compiles (gravity2_b200/csrc/classify_kernels.cuh, philox_kernels.cuh).  Row maxima
follow np.max / np.argmax (first maximum; a NaN wins, the first NaN is the arg-maximum: what
app/utils/classification_utils.py:63-68 gets from numpy); Philox4x32-10 is held to the Random123 restatement."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(emul_build):
    return emul_build("small_emul")


@pytest.mark.parametrize("rows,ld,r0,r1,c0,c1", [(5, 7, 0, 5, 0, 7), (9, 700, 2, 9, 13, 650), (6, 300, 1, 4, 255, 256)])
def test_emulated_rowmax_block_equals_numpy(emul, rows, ld, r0, r1, c0, c1):
    rng = np.random.default_rng(rows * ld)
    S = np.round(rng.random((rows, ld)), 2)                       # two decimals: plenty of tied maxima
    S[r0, c0:c1] = 0.5                                            # a constant row: the first column wins
    if r1 - r0 > 2:
        S[r0 + 1, (c0 + c1) // 2] = np.nan                        # NaN wins ...
        S[r0 + 1, c1 - 1] = np.nan                                # ... and the first NaN is the arg-maximum
        S[r0 + 2, c0:c1] = -np.inf
    out = subprocess.run([emul, "rowmax"] + [str(v) for v in (rows, ld, r0, r1, c0, c1)], input=S.tobytes(), capture_output=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    k = r1 - r0
    mx = np.frombuffer(out.stdout[:8 * k], dtype=np.float64)
    arg = np.frombuffer(out.stdout[8 * k:], dtype=np.int64)
    blk = S[r0:r1, c0:c1]
    assert np.array_equal(mx, np.max(blk, axis=1), equal_nan=True)
    assert np.array_equal(arg, np.argmax(blk, axis=1))


def test_emulated_gather_cells(emul):
    rng = np.random.default_rng(2)
    S = rng.random((40, 53))
    ii, jj = rng.integers(0, 40, 777), rng.integers(0, 53, 777)
    out = subprocess.run([emul, "gather", "40", "53", "777"], input=S.tobytes() + ii.astype(np.int64).tobytes() + jj.astype(np.int64).tobytes(),
                         capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert np.array_equal(np.frombuffer(out.stdout, dtype=np.float64), S[ii, jj])


@pytest.mark.parametrize("seed,b0,nb,p", [(0, 0, 1, 4), (20241019, 0, 3, 1031), (0xDEADBEEF12345678, 5, 2, 2050)])
def test_emulated_philox_kernel_equals_random123_restatement(emul, seed, b0, nb, p):
    from oracle import philox_oracle as P
    out = subprocess.run([emul, "philox", str(seed), str(b0), str(nb), str(p)], capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr
    idx = np.frombuffer(out.stdout[:8 * nb * p], dtype=np.int64).reshape(nb, p)
    w = np.frombuffer(out.stdout[8 * nb * p:], dtype=np.int32).reshape(nb, p)
    ref = P.philox_indices(seed, b0 + nb, p)[b0:]                 # replicate b0 + b of the stream
    assert np.array_equal(idx, ref)
    assert np.array_equal(w, np.stack([np.bincount(r, minlength=p) for r in ref]))
    if seed == 0:
        assert P.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
