"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("ONEPROT_CPU_SHIM") in ("1", "emu", "lib"):      # tests/cpu_shim.py: GPU test files on CPU stand-ins (test hygiene only)
        from tests import cpu_shim
        cpu_shim.install()
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        # a device-side hang must end the test process instead of wedging the box: the kernels trap after
        # ONEPROT_WAIT_TRAP_CYCLES on their own, this bounds everything else (collectives, host waits)
        try:
            import pytest_timeout  # noqa: F401
            for item in items:
                if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                    item.add_marker(pytest.mark.timeout(900, method="thread"))
        except ImportError:
            pass
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
