import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

os.environ.setdefault("SOLVERZ_B200_CACHE", str(ROOT / "build" / "model_cache"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def libszb():
    from solverz_b200.num_api import lib
    if not lib.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build_runtime()
    return lib.load()
