"""The presence / absence kernels, run on the CPU: tests/emul/popc_emul.cpp includes the SAME source
the GPU build compiles (gravity2_b200/csrc/popc_kernels.cuh) and runs it on host threads with the launch shapes of
popc_kernels.cu -- bit packing by warp ballot, the triangular tile enumeration, the three-stage ring, ragged last tiles and
column blocks, the mirror through shared memory, the packed-tile form of the multi-GPU path and its unpacking.  Checked bit
for bit against the oracle's restatement of shared_pphmm_ratio (app/utils/shared_pphmm_graphs.py:59-77); the GPU tests check
the compiled kernels."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(emul_build):
    return emul_build("popc_emul")


@pytest.mark.parametrize("packed", [0, 1])
@pytest.mark.parametrize("n,p,density", [(1, 1, 1.0), (5, 63, 0.5), (130, 1025, 0.1), (257, 2100, 0.03), (128, 1024, 0.2)])
def test_emulated_presence_kernels_equal_oracle(emul, n, p, density, packed):
    from oracle import gravity_oracle as O
    rng = np.random.default_rng(n * 1000 + p)
    tab = np.where(rng.random((n, p)) < density, np.round(rng.gamma(2.0, 60.0, (n, p)), 1) + 0.1, 0.0)
    if n > 4:
        tab[3] = 0.0                                  # an empty genome
        tab[2, 0] = np.nan                            # NaN != 0 counts as present, as in numpy
        tab[4] = tab[1]                               # a duplicate
    out = subprocess.run([emul, str(n), str(p), str(packed)], input=np.ascontiguousarray(tab).tobytes(), capture_output=True,
                         timeout=900)
    assert out.returncode == 0, out.stderr
    buf = np.frombuffer(out.stdout, dtype=np.int32)
    inter, nnz = buf[:n * n].reshape(n, n), buf[n * n:]
    present = (tab != 0)
    assert np.array_equal(nnz, present.sum(1))
    ref = present.astype(np.int32) @ present.astype(np.int32).T
    assert np.array_equal(inter, ref)
    if n <= 130:                                      # the literal restatement of the reference's triple loop (slow)
        assert np.array_equal(inter, np.array(O.shared_pphmm_ratio(tab)))
