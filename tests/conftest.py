import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand; the prebuilt .so travels to the GPU box."""
    import oc_oracle
    return oc_oracle.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own object code (oracle/_ref). Present in the build container and, prebuilt, on the GPU box."""
    import oc_oracle
    if not oc_oracle.Reference.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return oc_oracle.Reference()


@pytest.fixture(scope="session")
def config1():
    return np.load(os.path.join(GOLDEN, "config1_features.npz"))


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def built():
    """Make sure both product libraries exist (nvcc cross-compiles without a GPU)."""
    from opencalibration_b200 import build
    return build.build_all()


@pytest.fixture(scope="session")
def gpu(built):
    """Initialised C ABI on cuda:0. GPU tests fail (not skip) when the CUDA library cannot run."""
    from opencalibration_b200 import capi
    capi.init(0)
    return capi


@pytest.fixture(scope="session")
def hostlib(built):
    from opencalibration_b200 import host
    host.lib()
    return host
