import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle and the product library exist (built in-tree, see __graft_entry__.build)."""
    import __graft_entry__ as g

    if not (os.path.exists(os.path.join(REPO, "oracle", "liboracle.so"))
            and os.path.exists(os.path.join(REPO, "atlas_b200", "libsptrans_b200.so"))):
        g.build()
