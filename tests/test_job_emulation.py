"""The PL1 bootstrap loop body on the kernels' own source, without a GPU, against the UNMODIFIED reference:
for the golden replicates of oracle/make_golden.py (the reference's np.random.choice draws, app/src/pl1_graphs.py:78-107) the
three emulated stages are chained as the bootstrap job chains them (api.cu, replicates_impl) --
  gom_emul   GOM signature tables of every replicate from column multiplicities (location table = signature table: quirk Q3),
  bs_emul    exact PPHMM pair sums of every replicate from the pair lists; a replicate's row sums are its diagonal pair sums
             (or mma_emul: the same sums from the tensor-core kernel over the tcgen05 model, the dense tables' path),
  gj_emul    job_tail: fixed-point image of the GOM tables, then the fused tail (GOM Jaccard + combine + 1 - S + mirror) --
and the distance matrices are compared with the reference's (1e-6, the north_star tolerance)."""
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import gravity_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exes(emul_build):
    return {name: emul_build(name) for name in ("gom_emul", "bs_emul", "gj_emul", "mma_emul")}


def run(exe, args, blob):
    r = subprocess.run([exe] + [str(a) for a in args], input=blob, capture_output=True, timeout=1800)
    assert r.returncode == 0, (exe, r.returncode, r.stderr)
    return np.frombuffer(r.stdout, dtype=np.float64)


def ws_index(r, c):
    return ((((r >> 4) << 3) + (c >> 4)) << 8) + ((r & 15) << 4) + (c & 15)


def emulated_pl1(exes, sig, taxo, ids, idx, tensor_core=False):
    """Distance matrices (nb, n, n) of the PL1 replicates with resample indices idx (nb, p), stage by stage as the job runs them.
    tensor_core: the pair sums come from the tcgen05 kernel (the path the job's cost model picks for dense tables) instead of
    the pair-list kernels."""
    n, p = sig.shape
    G, nb = len(ids), len(idx)
    w = np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32)
    # GOM tables of the replicates (PL1: the GOM database is built from the signature table's rows, pl1_graphs.py:83-94)
    db = O.gomdb(taxo, sig, ids)
    rows = [np.asarray(db[k], dtype=np.float64).reshape(-1, p) for k in ids]
    offs = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    gom = run(exes["gom_emul"], [n, p, G, nb, 0, 1], sig.tobytes() + offs.tobytes() + np.concatenate(rows).tobytes() + w.tobytes())
    # exact PPHMM pair sums of the replicates
    nt = (n + 127) // 128
    ntl = nt * (nt + 1) // 2
    if tensor_core:
        buf = run(exes["mma_emul"], [n, p, nb], sig.tobytes() + w.tobytes())
        ws = buf[2 + nb * n:].reshape(nb, ntl, 128 * 128)
    else:
        buf = run(exes["bs_emul"], [n, p, nb], sig.tobytes() + w.tobytes())
        ws = buf[1 + n:].reshape(nb, ntl, 128 * 128)
    tile_of = {(bi, bj): t for t, (bi, bj) in enumerate((a, b) for a in range(nt) for b in range(a, nt))}
    prs = np.array([[ws[b, tile_of[(i // 128, i // 128)], ws_index(i % 128, i % 128)] for i in range(n)] for b in range(nb)])
    ws = np.where(np.isnan(ws), 0.0, ws)                           # (blocks below the diagonal are never written, never read)
    # the fused tail, as the job runs it (p_scale: api.cu gv2_job_create)
    p_scale = math.ldexp(1.0, -math.frexp(sig.max() * 255.0 * p)[1])
    return run(exes["gj_emul"], ["job_tail", n, G, nb, repr(p_scale), 1.0, 1], ws.tobytes() + prs.tobytes() + gom.tobytes()).reshape(nb, n, n)


@pytest.mark.parametrize("tensor_core", [False, True])
@pytest.mark.parametrize("tag", ["A", "B"])
def test_emulated_pl1_replicates_equal_reference_golden(exes, tag, tensor_core):
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    sig, taxo, ids = g[f"{tag}_sig"], g[f"{tag}_taxo"], [str(x) for x in g[f"{tag}_gom_ids"]]
    idx, D_ref = g[f"{tag}_pl1_idx"], g[f"{tag}_pl1_D"]
    D = emulated_pl1(exes, sig, taxo, ids, idx, tensor_core)
    assert D.shape == D_ref.shape
    err = np.abs(D - D_ref)
    assert err.max() <= 1e-6, err.max()
    for b in range(len(idx)):
        assert np.array_equal(D[b], D[b].T)


@pytest.mark.parametrize("tensor_core", [False, True])
@pytest.mark.parametrize("tag", ["A", "B"])
def test_emulated_pl2_replicates_equal_reference_golden(exes, tag, tensor_core):
    """The PL2 loop body (row a9, app/src/virus_classification.py:290-315): sorted resample indices (multiplicities do not see
    the order), the real location table, the GOM database from the reference rows only, similarity (not distance) output."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    sig, loc, taxo, ids = g[f"{tag}_sig"], g[f"{tag}_loc"], g[f"{tag}_taxo"], [str(x) for x in g[f"{tag}_gom_ids"]]
    idx, S_ref, n_ref = g[f"{tag}_pl2_idx"], g[f"{tag}_pl2_S"], int(g[f"{tag}_pl2_nref"])
    n, p = sig.shape
    G, nb = len(ids), len(idx)
    w = np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32)
    db = O.gomdb(taxo[:n_ref], loc[:n_ref], ids)
    rows = [np.asarray(db[k], dtype=np.float64).reshape(-1, p) for k in ids]
    offs = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    gom = run(exes["gom_emul"], [n, p, G, nb, 0, 1], loc.tobytes() + offs.tobytes() + np.concatenate(rows).tobytes() + w.tobytes())
    if tensor_core:
        ws = run(exes["mma_emul"], [n, p, nb], sig.tobytes() + w.tobytes())[2 + nb * n:].reshape(nb, 1, 128 * 128)
    else:
        ws = run(exes["bs_emul"], [n, p, nb], sig.tobytes() + w.tobytes())[1 + n:].reshape(nb, 1, 128 * 128)
    prs = np.array([[ws[b, 0, ws_index(i, i)] for i in range(n)] for b in range(nb)])
    ws = np.where(np.isnan(ws), 0.0, ws)
    p_scale = math.ldexp(1.0, -math.frexp(sig.max() * 255.0 * p)[1])
    S = run(exes["gj_emul"], ["job_tail", n, G, nb, repr(p_scale), 1.0, 0], ws.tobytes() + prs.tobytes() + gom.tobytes()).reshape(nb, n, n)
    err = np.abs(S - S_ref)
    assert err.max() <= 1e-6 * max(1.0, np.abs(S_ref).max()), err.max()
