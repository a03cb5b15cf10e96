import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools", "synth"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the real reference at /root/reference")


def pytest_collection_modifyitems(config, items):
    try:
        import refharness
        have_ref = refharness.available()
    except Exception:  # noqa: BLE001
        have_ref = False
    skip_ref = pytest.mark.skip(reason="reference flac-py not mounted")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
