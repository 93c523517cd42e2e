import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference sources (/root/reference, or oracle/_ref made by oracle/build_ref.py)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_import            # mounted tree, $DPDFNET_REFERENCE, or the verbatim copy in oracle/_ref
    have_ref = ref_import.available()
    skip_ref = pytest.mark.skip(reason="reference sources not available (run oracle/build_ref.py where /root/reference is mounted)")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
