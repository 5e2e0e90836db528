import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def l1():
    """FP64 CPU restatement (oracle level 1); built on demand by oracle/Makefile."""
    import oracle_lib
    return oracle_lib.L1()


@pytest.fixture(scope="session")
def l0():
    """The unmodified reference behind oracle/l0_harness.cpp; only where oracle/_ref was built."""
    import oracle_lib
    if not oracle_lib.L0.available():
        pytest.skip("oracle/_ref/libvpt_l0.so not present (reference tree absent when oracle was built)")
    lib = oracle_lib.L0()
    lib.reset_scene(); lib.set_quirks(3)
    return lib


@pytest.fixture(scope="session")
def units():
    return dict(np.load(os.path.join(GOLDEN, "units.npz")))


@pytest.fixture(scope="session")
def paths():
    return dict(np.load(os.path.join(GOLDEN, "paths.npz")))


@pytest.fixture(scope="session")
def vpt():
    """The product's Python host layer with the CUDA library loaded (built on demand, in-tree)."""
    import minimal_volumetric_path_tracer_b200 as v
    from minimal_volumetric_path_tracer_b200 import build
    build.build_all()
    v.load_library()
    return v


@pytest.fixture(scope="session")
def gpu(vpt):
    """GPU tests must run the CUDA path: fail loudly (do not skip) when no device is usable."""
    n = vpt.device_count()
    assert n > 0, "no CUDA device visible: -m gpu tests must run on a GPU box"
    return vpt


def rel_err(got, want, floor=0.0):
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.maximum(np.abs(want), floor if floor > 0 else np.finfo(np.float64).tiny)


def vec_rel_err(got, want):
    """norm-wise relative error of row vectors"""
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    return np.linalg.norm(got - want, axis=-1) / np.maximum(np.linalg.norm(want, axis=-1), 1e-300)


def pytest_collection_modifyitems(config, items):
    """every test gets a hard time limit (pytest-timeout): a hung kernel must not eat the GPU lease"""
    for item in items:
        if item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(240 if item.get_closest_marker("gpu") else 600))
