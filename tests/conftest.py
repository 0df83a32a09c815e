import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def g():
    """The product package, with the library built (nvcc cross-compiles here without a GPU)."""
    import ghicp_b200
    ghicp_b200.build_library()
    return ghicp_b200


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture()
def scratch_cwd(tmp_path, monkeypatch):
    """The reference's Km::output writes Corres.txt into the CWD (src/km.cpp:148)."""
    monkeypatch.chdir(tmp_path)
    return tmp_path
