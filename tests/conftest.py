import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", category=UserWarning)
warnings.filterwarnings("ignore", category=FutureWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box via gpurun)")


def pytest_sessionstart(session):
    """Development builds only (-DM4D_DEV): M4D_DEV_FLAGS=0x122 runs the suite against a kernel variant."""
    flags = os.environ.get("M4D_DEV_FLAGS")
    if flags:
        from more4d_b200 import _lib
        _lib.dev_set_flags(int(flags, 0))


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden():
    from safetensors.torch import load_file

    def _load(name):
        return load_file(os.path.join(GOLDEN, name + ".safetensors"))
    return _load
