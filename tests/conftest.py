import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden", "reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    """(name, k, input text bytes, expected dump lines) for the reference's 9 graph-build goldens (all k=3)."""
    out = []
    for t in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.test.txt"))):
        name = os.path.basename(t)[: -len(".test.txt")]
        with open(t, "rb") as f:
            text = f.read()
        with open(t.replace(".test.txt", ".expected.txt")) as f:
            expected = f.read().splitlines()
        out.append((name, 3, text, expected))
    return out


@pytest.fixture(scope="session")
def goldens():
    return golden_cases()
