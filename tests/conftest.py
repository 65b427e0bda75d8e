import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(1, os.path.join(ROOT, "tests"))      # tests/matrices.py, tests/components.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle, build_oracle
    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference objects; skipped where oracle/_ref was never built."""
    from oracle.bindings import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/librefapi.so not built (needs /root/reference)")
    return Reference()


@pytest.fixture(scope="session")
def tsp():
    import travellingsalesmanoptimization_b200 as T
    if T.device_count() <= 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box (there is no CPU fallback)")
    return T
