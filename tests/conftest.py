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


@pytest.fixture(scope="session")
def emu_harness_path(tmp_path_factory):
    """tests/harness/kernel_logic_harness.cpp: the product's .cu files compiled as plain C++ against the host emulation shim
    (tests/harness/cuda_emu), built once per session.  GHICP_EMU_CXXFLAGS="-fsanitize=address,undefined
    -fno-omit-frame-pointer" (with LD_PRELOAD=libasan.so) runs the emulated kernels under the sanitizers."""
    import subprocess
    out = tmp_path_factory.mktemp("emu") / "libkernel_logic_harness.so"
    src = os.path.join(ROOT, "tests", "harness", "kernel_logic_harness.cpp")
    extra = os.environ.get("GHICP_EMU_CXXFLAGS", "").split()
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-DGHICP_EMU_HOST"] + extra +
                       ["-I" + os.path.join(ROOT, "tests", "harness", "cuda_emu"), "-x", "c++", "-shared", "-o", str(out), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return str(out)
