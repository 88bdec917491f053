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
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout)")


def pytest_collection_modifyitems(config, items):
    """A hung kernel must fail one test, not eat the whole GPU visit: every GPU test gets a time limit (pytest-timeout, if
    installed; the full-size and multi-rank tests get more)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if "gpu" in item.keywords and not any(m.name == "timeout" for m in item.iter_markers()):
            big = "fullsize" in item.nodeid or "multirank" in item.nodeid
            item.add_marker(pytest.mark.timeout(900 if big else 240))


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
