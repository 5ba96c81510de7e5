import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def hb():
    import hexl_b200

    hexl_b200.lib()
    return hexl_b200


@pytest.fixture(scope="session")
def acquired(hb):
    """Mirror of the reference's gtest Environment (tests/fpga_context.h:8-14)."""
    hb.acquire_FPGA_resources()
    yield hb
    hb.release_FPGA_resources()


def set_variant(hb, name, value):
    """set_option for kernel variants that only exist in builds with `make EXPERIMENTAL=1` (they measured
    slower than the defaults): skips the test when the variant is compiled out."""
    try:
        hb.set_option(name, value)
    except hb.HexlB200Error as e:
        if "EXPERIMENTAL" in str(e):
            pytest.skip(f"{name}={value} is compiled out of the default build")
        raise
