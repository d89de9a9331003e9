import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    p = graft.load_package()
    missing = [q for q in (p.LIB_PATH, p.ORACLE_LIB_PATH, p.SYNTH_LIB_PATH) if not q.exists()]
    if missing:
        graft.build()
    return p


@pytest.fixture(scope="session")
def engine(pkg):
    return pkg.load_engine()


@pytest.fixture(scope="session")
def oracle(pkg):
    return pkg.load_oracle()


@pytest.fixture(scope="session")
def reference(pkg):
    if not pkg.REF_LIB_PATH.exists():
        pytest.skip("oracle/_ref/libhprlp_ref.so not built (needs /root/reference at build time)")
    return pkg.load_reference()
