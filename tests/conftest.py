import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _yardstick_default():
    """The product's only precision is 'kernel'.  The parity tests compare it (and the CPU oracle) with the torch-ops
    yardstick (yardstick/torch_modes.py, test infrastructure): installed here, and 'fp32' -- the reference's own op
    sequence -- is what a test sees unless it selects 'kernel' itself.  The product default is restored afterwards."""
    import yardstick
    from cooperative_training_and_latent_space_data_augmentation_b200 import conv_blocks
    yardstick.install()
    conv_blocks.set_precision("fp32")
    yield
    conv_blocks.set_precision("kernel")
