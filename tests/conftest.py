import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on the CPU")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run through gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """make sure the oracle (test infrastructure) and the product library exist"""
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")) or \
       not os.path.exists(os.path.join(ROOT, "oracle", "oracle_cli")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so", "oracle_cli"])
    if not os.path.exists(os.path.join(ROOT, "strainer2_b200", "libstrainer2_b200.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "strainer2_b200", "csrc")])
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden", "cases")


@pytest.fixture(scope="session")
def ref_dir():
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(d, "kmer_scrub_count")):
        pytest.skip("oracle/_ref (the compiled reference) is not present")
    return d
