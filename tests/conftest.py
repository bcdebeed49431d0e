import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """On a host without a CUDA device (this build container) a plain `pytest tests` SKIPS the gpu-marked tests instead
    of failing them; with a device present they run and the product fails loudly if its CUDA library is missing."""
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (pytest -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def imhd():
    """The product package (hyphenated directory name -> importlib).  Built artefacts are git-ignored, so a fresh
    checkout compiles the library first (nvcc cross-compiles sm_100a without a GPU)."""
    lib = os.path.join(ROOT, "imhd-cuda_b200", "libimhd_b200.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "imhd-cuda_b200", "csrc")])
    return importlib.import_module("imhd-cuda_b200")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as o

    need_ref = os.path.isdir("/root/reference/lib/on-device") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libimhd_ref_cpu.so"))
    if not os.path.exists(os.path.join(ROOT, "oracle", "libimhd_oracle.so")) or need_ref:
        o.build()
    return o


@pytest.fixture(scope="session")
def O(oracle_mod):
    return oracle_mod.Oracle()


@pytest.fixture(scope="session")
def R(oracle_mod):
    try:
        return oracle_mod.Reference()
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref/libimhd_ref_cpu.so not built (needs /root/reference)")


BOUNDS = (-3.14159, 3.14159) * 3  # build/on-device/input.inp:8-13


def make_case(O, oracle_mod, Nx, Ny, Nz, ic="screwpinch"):
    """Grids, spacings and initial state of a synthetic case, from the CPU oracle."""
    g = O.init_grids(BOUNDS, Nx, Ny, Nz)
    d = tuple(float(oracle_mod.grid_spacing(BOUNDS[2 * a], BOUNDS[2 * a + 1], n)) for a, n in enumerate((Nx, Ny, Nz)))
    Q = O.screwpinch_stride(1.0, *g) if ic == "screwpinch" else O.cubic_bennett_vortex_m0(2.0, 0.5, *g)
    return g, d, Q


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
