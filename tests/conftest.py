import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure only)."""
    from kestrel_b200 import capi
    path = os.path.join(ROOT, "oracle", "libkestrel_oracle.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])  # no-op when up to date
    return capi.Library(path, "kor_")


@pytest.fixture(scope="session")
def oracle_fma_lib():
    from kestrel_b200 import capi
    path = os.path.join(ROOT, "oracle", "libkestrel_oracle_fma.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])  # no-op when up to date
    return capi.Library(path, "kor_")


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; fails loudly when it has not been built."""
    from kestrel_b200 import capi
    return capi.load_gpu()
