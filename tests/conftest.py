import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def workloads():
    from iq_tool_b200 import baseline_workloads
    return baseline_workloads()


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    gdir = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gdir, "golden.json")) as f:
        meta = json.load(f)
    data = {}
    for name in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
        data[name] = dict(np.load(os.path.join(gdir, f"{name}.npz")))
    return meta, data


@pytest.fixture(scope="session")
def gpu():
    """The CUDA chain binding; skips (never falls back) when no device is present."""
    from iq_tool_b200 import gpu as g
    if g.device_count() < 1:
        pytest.skip("no CUDA device")
    return g
