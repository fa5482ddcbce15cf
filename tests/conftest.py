import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def emul_build(tmp_path_factory):
    """build(name) -> path of tests/emul/<name>.cpp compiled with g++ (a kernel family's own source on host threads); each
    program is built once per session, whichever test modules ask for it."""
    import shutil
    import subprocess
    built = {}

    def build(name):
        if shutil.which("g++") is None:
            pytest.skip("no host compiler")
        if name not in built:
            exe = str(tmp_path_factory.mktemp("emul") / name)
            subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-ffp-contract=off", "-o", exe,
                            os.path.join(REPO, "tests", "emul", name + ".cpp")], check=True)
            built[name] = exe
        return built[name]
    return build


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(REPO, "tests", "golden", "reference_golden.npz")
    return dict(np.load(path, allow_pickle=False))


@pytest.fixture(scope="session")
def golden_nbhd():
    """PphmmNeighbourhoodWeight != 0 outputs of the unmodified reference (oracle/make_golden.py neighbourhood)."""
    path = os.path.join(REPO, "tests", "golden", "reference_golden_neighbourhood.npz")
    return dict(np.load(path, allow_pickle=False))


def write_reference_csvs(tmp, sig, naive_loc):
    """The two CSVs SimilarityMat_Constructor reads, laid out as app/src/pl1_graphs.py:307-319 writes them."""
    import pandas as pd
    n, p = sig.shape
    names = [f"v{i}" for i in range(n)]
    cols = [f"PPHMM {c}" for c in range(p)]
    df = pd.DataFrame(sig, columns=cols)
    df.insert(0, "Virus name", names)
    f_sig = os.path.join(tmp, "sigs.csv")
    df.to_csv(f_sig)
    dl = pd.DataFrame(np.abs(naive_loc), columns=cols)
    dl.insert(0, "Virus name", names)
    f_loc = os.path.join(tmp, "locs.csv")
    dl.to_csv(f_loc, index=False)
    return {"PphmmLocs": f_loc, "PphmmAndGomSigs": f_sig}
