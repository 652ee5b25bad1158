import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _build_oracle_port():
    """The C restatement under oracle/ is a checker; build it once per session if missing."""
    import subprocess
    so = ROOT / "oracle" / "_build" / "libmasic_oracle.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "port"], check=True, capture_output=True)
