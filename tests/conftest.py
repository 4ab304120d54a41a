import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import json
    cases = {}
    for name in ("reference_golden.json", "reference_golden_r2.json"):   # r2: oracle/make_golden_r2.py (BASELINE sizes, new rows)
        with open(os.path.join(ROOT, "tests", "golden", name)) as f:
            cases.update(json.load(f)["cases"])
    return cases


@pytest.fixture(autouse=True)
def _poisoned_cuda_cache(request):
    """GPU tests start with NaN in every cached free block (see _util.poison_cuda_cache)."""
    if request.node.get_closest_marker("gpu") is not None:
        from _util import poison_cuda_cache
        poison_cuda_cache()
    yield
