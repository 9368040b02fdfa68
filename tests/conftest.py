import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (oracle/) and the product library if they are missing (CPU-only, no GPU needed)."""
    from oracle import cpu_oracle
    if not cpu_oracle.available("flip", "port"):
        cpu_oracle.build()
    import simsandgames_b200 as pkg
    if not os.path.exists(pkg.lib_path()):
        pkg.build_library()
    yield


@pytest.fixture(scope="session")
def has_ref():
    from oracle import cpu_oracle
    return cpu_oracle.available("flip", "ref") and cpu_oracle.available("euler", "ref")
