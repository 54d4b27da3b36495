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
