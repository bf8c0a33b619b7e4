"""pytest configuration: marker registration and shared fixtures.

``-m "not gpu"`` runs here (no GPU): oracle vs. golden vectors / reference, host
logic, C-ABI symbol checks.  ``-m gpu`` runs on a B200 and is the parity suite
proper (CUDA path vs. oracle through the C ABI).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Stable small argsort + libm transcendentals in NumPy (see oracle/ref_loader.py).
# Must happen before NumPy is first imported by the test process.
from oracle import ref_loader  # noqa: E402

# The pin only works if NumPy has not been imported yet (pytest.ini disables the
# plugins that would import it first).  Tests that compare NumPy-oracle floats
# bit-for-bit fall back to a 1e-6 tolerance when the pin could not take effect.
os.environ["MGD_NUMPY_PIN_EFFECTIVE"] = "0" if "numpy" in sys.modules else "1"
ref_loader.pin_numpy_env()

import pytest  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.build()
    return co
