import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    arrays = np.load(os.path.join(here, "golden.npz"))
    with open(os.path.join(here, "golden.json")) as fh:
        meta = json.load(fh)
    return arrays, meta


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    """The reference-built checker, or None where it was not built/shipped."""
    import oracle
    return oracle.reference()


@pytest.fixture(scope="session")
def checker(port, ref):
    """Best available CPU checker: the unmodified reference when present, else the pinned port."""
    return ref if ref is not None else port


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from phys_autodiff_b200 import ops
    c = ops.Context()
    yield c
    c.close()
