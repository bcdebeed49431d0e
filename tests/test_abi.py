"""CPU tests of the drop-in boundary: libimhd_b200.so loads, exports every symbol include/imhd_b200.h
declares, the ctypes table covers the same set, and compute entry points fail LOUDLY (no fallback)
when no CUDA device is usable."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "imhd_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(imhd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = header_functions()
    for need in ("imhd_predictor", "imhd_corrector", "imhd_fluid_bcs", "imhd_step_fused", "imhd_qint_plane",
                 "imhd_create", "imhd_ctx_step", "imhd_run_host", "imhd_init_screwpinch_stride"):
        assert need in names


def test_library_exports_every_declared_symbol(imhd):
    lib = imhd._lib.load()
    out = subprocess.run(["nm", "-D", "--defined-only", imhd._lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (imhd_[a-z0-9_]+)", out))
    declared = set(header_functions())
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert declared == set(imhd._lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.imhd_abi_version() == 1


def test_sm100a_only(imhd):
    out = subprocess.run(["cuobjdump", "--list-elf", imhd._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_scalar_host_helper(imhd):
    f = imhd.ops.wall_energy_fixed_point
    assert f(0.0, 100) == 0.0
    e = f(0.7312345, 100)
    assert f(e, 1) == e and abs(e - 0.7312345) < 1e-6


def test_no_silent_fallback_without_gpu(imhd):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for GPU-less hosts")
    lib = imhd._lib.load()
    assert not lib.imhd_create(16, 16, 16, 0)
    msg = lib.imhd_last_error().decode()
    assert "no usable CUDA device" in msg and "no CPU path" in msg
    with pytest.raises(imhd.ImhdError):
        imhd.Context(16, 16, 16)
    # a stateless operator on garbage pointers must return a CUDA error, not pretend to work
    rc = lib.imhd_predictor(None, None, 0, 0.0, 1e-4, 0.1, 0.1, 0.1, 16, 16, 16, None)
    assert rc != 0 and "GPUassert" in lib.imhd_last_error().decode()


def test_bad_arguments_are_rejected(imhd):
    lib = imhd._lib.load()
    assert lib.imhd_predictor(None, None, 0, 0.0, 1e-4, 0.1, 0.1, 0.1, 2, 16, 16, None) == imhd._lib.E_INVALID
    s = imhd.ops.make_slab(16, 16, 16, 0, 0.0, 1e-4, 0.1, 0.1, 0.1, k0=4, nzl=4, ghosts=0)
    assert lib.imhd_step_fused(None, None, None, None, None, C.byref(s), None) == imhd._lib.E_INVALID
    assert "ghosts" in lib.imhd_last_error().decode()


# ---- string-keyed registry (include/on-device/utils/configurers.hpp); pure host logic, no GPU needed ----------
def test_registry_lists_the_reference_keys(imhd):
    ops = imhd.ops
    assert ops.registry_names(ops.REG_INITIALIZER)[:3] == ["screwpinch", "screwpinch-stride", "cubic-bennett-vortex"]
    assert "fluidadvancelocal-nodiff" in ops.registry_names(ops.REG_CORRECTOR)
    assert {"corrector_advance-tp_nodiff", "corrector_advance-stride_nodiff"} <= set(ops.registry_names(ops.REG_PREDICTOR))
    assert ops.registry_names(ops.REG_FLUID_BCS) == ["pcrw-xy_pbc-z"]
    assert ops.registry_names(ops.REG_PREDICTOR_BCS) == ["pbc-z"]
    L = imhd._lib.load()
    assert L.imhd_registry_name(0, 99) is None and L.imhd_registry_count(17) == 0
    assert L.imhd_registry_initializer_nparams(b"screwpinch") == 2
    assert L.imhd_registry_initializer_nparams(b"orszag-tang") == -1


def test_registry_resolves_bundles_to_a_time_loop(imhd):
    ops = imhd.ops
    assert ops.registry_resolve_path("fluidadvancelocal-nodiff", "corrector_advance-tp_nodiff") == 0
    assert ops.registry_resolve_path("fluidadvancelocal-nodiff", "corrector_advance-stride_nodiff") == 0
    assert ops.registry_resolve_path("fluidadvancelocal", "corrector_advance-stride") == 1
    for args, msg in ((("nope", "corrector_advance-tp_nodiff"), "Unknown kernel bundle selected: nope"),
                      (("fluidadvancelocal-nodiff", "nope"), "Unknown I.V. kernel bundle selection: nope"),
                      (("fluidadvancelocal-nodiff", "corrector_advance-tp_nodiff", "all-pbc"), "Unknown bcs selected: all-pbc"),
                      (("fluidadvancelocal", "corrector_advance-tp_nodiff"), "different time loops")):
        with pytest.raises(imhd.ImhdError, match=msg):
            ops.registry_resolve_path(*args)
