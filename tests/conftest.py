import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """libkmpc.so is a build artefact (git-ignored): if the suite is started in a fresh checkout before
    __graft_entry__.build(), compile it once (nvcc cross-compiles sm_100a without a GPU, about a minute)."""
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        from koopman_mpc_portfolio_rebalancing_b200 import build as kb
        kb.build(verbose=False)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


def sd_from_npz(z):
    return {k[4:]: z[k] for k in z.files if k.startswith("sd::")}
