import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import binding

    return binding.load()


def make_oracle(**kw):
    from oracle.binding import Oracle

    return Oracle(**kw)


def make_renderer():
    """The product path.  No fallback: if the library or the GPU is missing this raises."""
    from ray_tracing_gallery_b200 import native

    return native.Renderer(0)
