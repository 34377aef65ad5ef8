import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run through gpurun)")


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from transception_b200 import ops
    lib = ops.load_library()          # raises (does not skip) if the .so is missing on a GPU box
    assert lib.tcx_device_ok() == 1, lib.tcx_last_error().decode()
    return lib
