import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """libpraga_b200.so, (re)built with nvcc when sources changed; on a box without nvcc the prebuilt .so is used."""
    from praga_b200 import build
    try:
        build.build()
    except RuntimeError:
        if not os.path.exists(build.LIB):
            raise
    return build.LIB


@pytest.fixture(scope="session")
def engine(built_lib):
    import praga_b200 as pb
    eng = pb.Engine(0)          # raises (no CPU fallback) when there is no B200
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def ref():
    """The reference oracle (oracle/_ref, built from /root/reference where present, else the prebuilt .so)."""
    from oracle import binding
    if os.path.isdir("/root/reference/agrolib"):
        binding.build("ref")
    if not binding.have_ref():
        pytest.skip("oracle/_ref/libpraga_ref.so not available")
    return binding.Oracle("reference")
