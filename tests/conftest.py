import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def emulated_abi(monkeypatch):
    """Swap the ctypes wrappers of gail_carla_b200._abi for their CPU statements (oracle/abi_emu.py).

    Test-only: lets the host-side classes run in a container without a GPU so their layout bookkeeping and
    hand-derived backward passes can be checked against oracle.ref_path and the golden vectors.
    """
    from gail_carla_b200 import _abi
    from oracle import abi_emu
    for name in dir(abi_emu):
        fn = getattr(abi_emu, name)
        if callable(fn) and not name.startswith("_") and hasattr(_abi, name) and name not in ("call", "load_library"):
            monkeypatch.setattr(_abi, name, fn)
    monkeypatch.setattr(_abi, "EMULATED", True, raising=False)
    yield abi_emu
