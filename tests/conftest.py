import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One rrt_ctx on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    from rs_ray_toy_b200.aggregate import Context

    c = Context(0)
    yield c
    c.close()
