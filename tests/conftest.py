import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the product library and the oracle once per session (nvcc/g++ work without a GPU)."""
    from diligentfx_b200 import build as product_build
    from oracle import oracle_py
    product_build.build()
    oracle_py.build()
    return True


@pytest.fixture(scope="session")
def seq_small():
    """4 consecutive synthetic frames, odd dimensions (hits the odd-size mip branches)."""
    from diligentfx_b200 import synth
    return synth.generate_sequence(333, 187, 4)


@pytest.fixture(scope="session")
def seq_even():
    from diligentfx_b200 import synth
    return synth.generate_sequence(256, 144, 4)
