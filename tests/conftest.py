"""Test configuration.  `-m "not gpu"` = oracle vs the reference's golden vectors, host logic, ABI
symbol/compile checks (no GPU needed); `-m gpu` = parity of the CUDA path with the oracle through the C ABI."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(directory):
    res = subprocess.run(["make", "-C", os.path.join(ROOT, directory), "-j8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (test infrastructure)."""
    from sqlrs_b200.host import ffi

    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        _make("oracle")
    return ffi.Library(path, "sqlrs_oracle_")


@pytest.fixture(scope="session")
def cuda_lib(request):
    """The product library; fails loudly when it is missing (no fallback)."""
    from sqlrs_b200.host import ffi

    if os.environ.get("SQLRS_TEST_HARNESS_SELFCHECK") == "1":  # CPU dry run of the GPU test harness itself (oracle vs oracle)
        return request.getfixturevalue("oracle")
    if not os.path.exists(ffi.library_path()):
        _make("sqlrs_b200/csrc")
    return ffi.load()


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def lib(request):
    """Every golden case runs against the oracle on CPU and, on the GPU box, against the CUDA library."""
    return request.getfixturevalue("oracle" if request.param == "oracle" else "cuda_lib")
