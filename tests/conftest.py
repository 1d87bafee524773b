import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Per-case parity statistics collected by the GPU tests (tests/parity.py:record) -> gpurun_out/parity_report.json."""
    try:
        from tests import parity
    except Exception:
        return
    if not parity.REPORT:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity_report.json")
    merged = {}
    try:                                   # several pytest sessions of one GPU call (full suite, sanitizer subsets) add up
        merged = json.load(open(path))
    except Exception:
        pass
    merged.update(parity.REPORT)
    with open(path, "w") as f:
        json.dump(merged, f, indent=1, sort_keys=True)
