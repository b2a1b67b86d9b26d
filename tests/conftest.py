import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The suite needs the built library (ABI tests here, every GPU test on the box) and the C oracle.  Both are
    # git-ignored build products: build them once when they are missing (nvcc / gcc cross-compile without a GPU).
    lib = os.path.join(ROOT, "jaxrenderer_b200", "lib", "libjr_b200.so")
    if not os.path.exists(lib) and not os.environ.get("JR_B200_LIB"):
        import __graft_entry__ as entry

        entry.build()


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
