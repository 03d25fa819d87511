import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "smith-waterman-fpga-module_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as om
    om.build_oracle()
    return om


@pytest.fixture(scope="session")
def pkg():
    """The product package (hyphenated directory name => importlib).  On a fresh checkout the
    native library is built first (nvcc cross-compiles sm_100a without a GPU)."""
    mod = importlib.import_module(PKG_NAME)
    if not os.path.exists(mod.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "bin", "sw_b200_cli")):
        import __graft_entry__
        __graft_entry__.build()
    return mod
