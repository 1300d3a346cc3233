import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def fixture_npz():
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "fixture_evidence.npz"))


@pytest.fixture(scope="session")
def fixture_batch(fixture_npz):
    from util import batch_from_npz
    return batch_from_npz(fixture_npz)


@pytest.fixture()
def oracle_scorer(oracle, monkeypatch):
    """CPU tests of the host plumbing: the parity oracle stands in for the CUDA engine.  It is installed by
    monkeypatching genotype.score (the product module has no injection seam) and reads the product's COMPACT
    batches through compact.wide_from_compact, so the encoding is exercised too."""
    from svtyper_b200 import compact as cp, genotype

    def scorer(batch, **params):
        return oracle.score(cp.wide_from_compact(batch), **params)
    monkeypatch.setattr(genotype, "score", scorer)
    yield scorer
