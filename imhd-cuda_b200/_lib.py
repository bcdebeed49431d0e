"""ctypes binding of libimhd_b200.so -- every symbol include/imhd_b200.h declares.

The library is the product; this file only declares signatures.  Loading fails loudly when the
shared object has not been built (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C imhd-cuda_b200/csrc``): there is no Python or CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# IMHD_B200_LIB: load another build of the same ABI (kernel experiments, tools/experiments/)
LIB_PATH = os.environ.get("IMHD_B200_LIB") or os.path.join(PKG_DIR, "libimhd_b200.so")

PATH_A, PATH_B = 0, 1
E_INVALID, E_STATE, E_IO = 10001, 10002, 10003


class Slab(C.Structure):
    """``imhd_slab`` of include/imhd_b200.h."""

    _fields_ = [
        ("Nx", C.c_int), ("Ny", C.c_int), ("Nz", C.c_int),
        ("k0", C.c_int), ("nzl", C.c_int), ("ghosts", C.c_int), ("path", C.c_int),
        ("D", C.c_float), ("dt", C.c_float), ("dx", C.c_float), ("dy", C.c_float), ("dz", C.c_float),
        ("corner_e", C.c_float),
    ]


class Stability(C.Structure):
    """``imhd_stability`` of include/imhd_b200.h."""

    _fields_ = [("max_lhs", C.c_float), ("i", C.c_int), ("j", C.c_int), ("k", C.c_int),
                ("violations", C.c_ulonglong), ("dt_new", C.c_float)]


_f, _i, _p, _u64 = C.c_float, C.c_int, C.c_void_p, C.c_uint64
_dims = [_i, _i, _i]
_op = [_p, _p, _i, _f, _f, _f, _f, _f] + _dims + [_p]

# name -> (restype, argtypes); the single source the ABI test checks against the header
SIGNATURES = {
    "imhd_abi_version": (_i, []),
    "imhd_last_error": (C.c_char_p, []),
    "imhd_launch_count": (_u64, []),
    "imhd_predictor": (_i, _op),
    "imhd_corrector": (_i, _op),
    "imhd_fluid_bcs": (_i, _op),
    "imhd_initial_bcs": (_i, [_p] + _dims + [_p]),
    "imhd_init_grids": (_i, [_p, _p, _p] + [_f] * 6 + _dims + [_p]),
    "imhd_init_screwpinch_stride": (_i, [_p, _f, _p, _p, _p] + _dims + [_p]),
    "imhd_init_cubic_bennett_vortex_m0": (_i, [_p, _f, _f, _p, _p, _p] + _dims + [_p]),
    "imhd_init_cubic_bennett_vortex": (_i, [_p, _p, _p, _p] + _dims + [_p]),
    "imhd_init_zpinch": (_i, [_p, _f, _p, _p, _p] + _dims + [_p]),
    "imhd_init_screwpinch": (_i, [_p, _f, _f, _p, _p, _p] + _dims + [_p]),
    "imhd_step_fused": (_i, [_p, _p, _p, _p, _p, C.POINTER(Slab), _p]),
    "imhd_step_fused_planes": (_i, [_p, _p, _p, _p, _p, C.POINTER(Slab), _i, _i, _p]),
    "imhd_step_fused_ends": (_i, [_p, _p, _p, _p, _p, C.POINTER(Slab), _i, _i, _i, _i, _p]),
    "imhd_stability_scan": (_i, [_p, C.POINTER(Slab), C.POINTER(Stability), _p]),
    "imhd_ctx_stability": (_i, [_p, _f, C.POINTER(Stability)]),
    "imhd_ctx_step_adaptive": (_i, [_p, _i, _i, _f, _f, _p]),
    "imhd_ctx_set_dt": (_i, [_p, _f]),
    "imhd_wall_energy_fixed_point": (_f, [_f, _i]),
    "imhd_set_chunk": (None, [_i]),
    "imhd_set_kernel_variant": (None, [_i]),
    "imhd_qint_plane": (_i, [_p, _p, _i, C.POINTER(Slab), _p]),
    "imhd_create": (_p, _dims + [_i]),
    "imhd_destroy": (None, [_p]),
    "imhd_ctx_init_grids": (_i, [_p] + [_f] * 6),
    "imhd_ctx_init_screwpinch_stride": (_i, [_p, _f]),
    "imhd_ctx_init_cubic_bennett_vortex_m0": (_i, [_p, _f, _f]),
    "imhd_registry_count": (_i, [_i]),
    "imhd_registry_name": (C.c_char_p, [_i, _i]),
    "imhd_registry_initializer_nparams": (_i, [C.c_char_p]),
    "imhd_ctx_initialize": (_i, [_p, C.c_char_p, _p, _i]),
    "imhd_registry_resolve_path": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(_i)]),
    "imhd_ctx_set_state": (_i, [_p, _p]),
    "imhd_ctx_set_spacing": (_i, [_p, _f, _f, _f]),
    "imhd_ctx_prime": (_i, [_p, _i, _f, _f]),
    "imhd_ctx_step": (_i, [_p, _i]),
    "imhd_ctx_step_granular": (_i, [_p, _i]),
    "imhd_ctx_get_state": (_i, [_p, _p]),
    "imhd_ctx_get_grids": (_i, [_p, _p, _p, _p]),
    "imhd_ctx_device_state": (_p, [_p]),
    "imhd_ctx_stream": (_p, [_p]),
    "imhd_ctx_synchronize": (_i, [_p]),
    "imhd_h5_write_fluidvars": (_i, [C.c_char_p, _p, _i, _i, _i, _i]),
    "imhd_h5_write_grid": (_i, [C.c_char_p, _p, _p, _p, _i, _i, _i]),
    "imhd_ctx_write_frame": (_i, [_p, C.c_char_p, _i]),
    "imhd_ctx_flush_output": (_i, [_p]),
    "imhd_ctx_write_grid": (_i, [_p, C.c_char_p]),
    "imhd_run_host": (_i, [_p, _p, _p, _i, _f, _f, _f, _f, _f, _i]),
    "imhd_create_multi": (_p, _dims + [_i, C.POINTER(_i)]),
    "imhd_create_slab": (_p, _dims + [_i, _i, _i, _p]),
    "imhd_nccl_unique_id": (_i, [_p, _i]),
    "imhd_ctx_num_slabs": (_i, [_p]),
    "imhd_ctx_slab_extent": (_i, [_p, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "imhd_ctx_set_state_local": (_i, [_p, _i, _p]),
    "imhd_ctx_get_state_local": (_i, [_p, _i, _p]),
    "imhd_set_edge_planes": (None, [_i]),
    "imhd_stability_mode": (None, [_i]),
    "imhd_fused_timing": (None, [_i]),
    "imhd_fused_timing_read": (_i, [C.POINTER(C.c_double), C.POINTER(_i), C.POINTER(C.c_longlong)]),
}


class ImhdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libimhd_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libimhd_b200.so and declare every signature.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not built: run `make -C {os.path.join(PKG_DIR, 'csrc')}` "
                "(there is no fallback implementation)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = res, args
        if os.environ.get("IMHD_KERNEL_VARIANT"):  # kernel experiments: see imhd_set_kernel_variant in the header
            lib.imhd_set_kernel_variant(int(os.environ["IMHD_KERNEL_VARIANT"]))
        _lib = lib
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise ImhdError(code, load().imhd_last_error().decode(errors="replace"))
