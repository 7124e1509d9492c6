import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own C++ (oracle/_ref).  Built here from /root/reference;
    on the GPU box the prebuilt library travels with the snapshot."""
    from oracle.oracle import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref/libsks_ref.so not built and /root/reference absent")
    return RefLib()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    return {name: np.load(os.path.join(g, name + ".npz"))
            for name in ("ref_general", "ref_torch", "kat_veri4pts", "ref_ge", "curand_mrg32k3a")}


@pytest.fixture(scope="session")
def sks():
    """The product library; building it needs nvcc only (no GPU)."""
    from sks_homography_b200 import build, lib
    build.build()
    return lib()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda:0")
