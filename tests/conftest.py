"""Shared fixtures. GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement of the reference path (oracle/limu_oracle.c)."""
    import oracle
    return oracle.load_port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled by oracle/Makefile (prebuilt binary; skip if absent)."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/liblimu_ref.so not built (needs /root/reference at build time)")
    return oracle.load_ref()


@pytest.fixture(scope="session")
def ref_mt():
    import oracle
    if not os.path.exists(oracle.REF_MT_SO):
        pytest.skip("oracle/_ref/liblimu_ref_mt.so not built")
    return oracle.load_ref(mt=True)


@pytest.fixture()
def rng():
    return np.random.default_rng(12345)


def random_pose(api, rng, trans=5.0, rot=0.5):
    x = rng.normal(size=6) * np.array([trans, trans, trans, rot, rot, rot])
    return api.se3_exp(x)
