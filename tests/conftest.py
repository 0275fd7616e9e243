import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load(name):
    with np.load(os.path.join(GOLDEN_DIR, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_unc():
    return _load("uncertainty.npz")


@pytest.fixture(scope="session")
def golden_agg():
    return _load("aggregation.npz")


@pytest.fixture(scope="session")
def golden_calib():
    return _load("calibration.npz")


@pytest.fixture(scope="session")
def golden_ncc_aurc():
    return _load("ncc_aurc.npz")


def case_names(d):
    return sorted({k.split("/")[0] for k in d})


@pytest.fixture(scope="session")
def golden_quantile():
    return _load("quantile.npz")


@pytest.fixture(scope="session")
def golden_ged_nll():
    return _load("ged_nll.npz")
