import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def fixtures_dir():
    return os.path.join(ROOT, "tests", "golden", "ref_fixtures")


@pytest.fixture(scope="session")
def ref_fasta(fixtures_dir):
    return os.path.join(fixtures_dir, "fasta.fa.gz")


@pytest.fixture(scope="session")
def ref_vcf(fixtures_dir):
    return os.path.join(fixtures_dir, "test.vcf.gz")


FULL_PARITY = "test_full_parity_gpu.py"


def pytest_collection_modifyitems(config, items):
    """The full-size parity module last: its oracle runs (minutes of single-thread CPU work) start with the session
    (fixture below) and finish beside the other GPU tests."""
    tail = [it for it in items if it.path.name == FULL_PARITY]
    if tail:
        items[:] = [it for it in items if it.path.name != FULL_PARITY] + tail


@pytest.fixture(scope="session", autouse=True)
def _full_parity_oracles_start_early(request):
    names = {it.name for it in request.session.items if it.path.name == FULL_PARITY}
    if names and len(request.session.items) > len(names):   # alone, the module starts them itself
        import test_full_parity_gpu as T
        T.JOBS.update(T.start_jobs(names))
    yield
