import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def emu_api():
    """The product sources compiled against the SIMT emulator (tests/emu): CPU-only logic check."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    import nim_snappy_b200 as s
    from nim_snappy_b200 import _lib
    return s.Snappy(_lib.SnbLibrary(os.path.join(ROOT, "tests", "emu", "libsnb_emu.so")))


@pytest.fixture(scope="session")
def gpu_api():
    """The real CUDA library through the C ABI; fails loudly when it is missing."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import nim_snappy_b200 as s
    from nim_snappy_b200 import _lib
    return s.Snappy(_lib.default())


def read_data(name):
    with open(os.path.join(DATA, name), "rb") as f:
        return f.read()


def read_golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


CORPUS = ["html", "urls.10K", "fireworks.jpeg", "paper-100k.pdf", "html_x_4", "alice29.txt", "asyoulik.txt",
          "lcet10.txt", "plrabn12.txt", "geo.protodata", "kppkn.gtb", "Mark.Twain-Tom.Sawyer.txt", "house.jpg"]
