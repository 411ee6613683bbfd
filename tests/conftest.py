import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure librtb200.so and liboracle.so exist (compiles them when missing)."""
    import __graft_entry__ as g
    from raytracinggpu_b200 import api
    from oracle import pyoracle
    if not os.path.exists(api.LIB_PATH) or not os.path.exists(pyoracle.ORACLE_SO):
        g.build()
    return True


@pytest.fixture(scope="session")
def cat_path(built):
    from oracle import pyoracle
    p = pyoracle.cat_obj_path()
    if p is None:
        pytest.skip("cat.obj not available (neither /root/reference nor oracle/_ref)")
    return p
