import glob
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "cardiac-ablation-ecm2_b200"))

GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases():
    return sorted(os.path.basename(f)[5:-4] for f in glob.glob(os.path.join(GOLDEN, "case_*.npz")))


def load_case(tag):
    d = dict(np.load(os.path.join(GOLDEN, f"case_{tag}.npz")))
    for k in ("p", "D1D", "Q1D", "NE", "ndofs"):
        d[k] = int(d[k][0])
    return d


@pytest.fixture(params=golden_cases())
def case(request):
    return load_case(request.param)


@pytest.fixture(scope="session")
def ctx():
    """One b200pa context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    import b200pa
    c = b200pa.Context(0)
    yield c
    c.close()
