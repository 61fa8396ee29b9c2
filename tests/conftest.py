import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# no network here: the drop-in VGG19 would otherwise try to download the pretrained weights like the reference does
# (models/vgg19_net.py:27) and raise; tests load seeded random weights themselves
os.environ.setdefault("FNST_VGG19_RANDOM_INIT", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
