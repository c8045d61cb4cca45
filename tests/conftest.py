import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def reference_fixtures():
    with open(os.path.join(GOLDEN, "reference_fixtures.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_env():
    """Installs the import shims (gymnasium / posggym stand-ins) for the oracle."""
    from oracle import shims

    have_ref = shims.install()
    return {"have_reference": have_ref}


def golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name))
