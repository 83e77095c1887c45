import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def R():
    """The unmodified reference (oracle/_ref/libroaring_ref.so) — checker only."""
    from oracle.refbind import ref
    return ref()


@pytest.fixture(scope="session")
def O():
    """The plain-C restatement (oracle/liboracle.so) — checker only."""
    from oracle.oraclebind import oracle
    return oracle()


@pytest.fixture(scope="session")
def rb():
    """The product (libroaring_b200.so through its Python mirror), initialised on cuda:0."""
    import croaring_b200 as m
    m.init(0)
    return m


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "realdata_golden.json")) as f:
        return json.load(f)
