"""Operator-level Python surface over the C ABI.

Two layers, both thin:

* free functions over torch CUDA tensors (``predictor``, ``corrector``, ``fluid_bcs``, ``step_fused`` ...):
  PyTorch supplies device memory and the current stream, the library does all the work.  They mirror
  the reference's kernel groups (argument order of include/on-device/*.cuh: state arrays, then
  ``D, dt, dx, dy, dz``; the grid size is taken from the tensor shape ``(8, Nz, Nx, Ny)``).
* ``Context`` -- the ``imhd_ctx`` the drop-in drivers use (host numpy in / out).

Nothing here computes: a missing or unloadable ``libimhd_b200.so`` raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PATH_A, PATH_B, Slab, check

VARS = ("rho", "rhovx", "rhovy", "rhovz", "Bx", "By", "Bz", "e")


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, shape=None):
    import torch

    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise TypeError("expected a contiguous float32 CUDA tensor")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return C.c_void_p(t.data_ptr())


def _dims(Q):
    if Q.dim() != 4 or Q.shape[0] != 8:
        raise ValueError("state tensors have shape (8, Nz, Nx, Ny)")
    _, Nz, Nx, Ny = Q.shape
    return Nx, Ny, Nz


def grid_spacing(lo: float, hi: float, n: int) -> float:
    """fp32 ``(hi - lo) / (n - 1)`` as the reference drivers compute dx (main.cu:98-100)."""
    return float(np.float32((np.float32(hi) - np.float32(lo)) / np.float32(n - 1)))


# ---- parity-granular operators ------------------------------------------------------------------
def predictor(Q, Qint, path, D, dt, dx, dy, dz):
    L = _lib.load()
    check(L.imhd_predictor(_dev(Q), _dev(Qint, Q.shape), path, D, dt, dx, dy, dz, *_dims(Q), _stream()))


def corrector(Q, Qint, path, D, dt, dx, dy, dz):
    L = _lib.load()
    check(L.imhd_corrector(_dev(Q), _dev(Qint, Q.shape), path, D, dt, dx, dy, dz, *_dims(Q), _stream()))


def fluid_bcs(Q, Qint, path, D, dt, dx, dy, dz):
    L = _lib.load()
    check(L.imhd_fluid_bcs(_dev(Q), _dev(Qint, Q.shape), path, D, dt, dx, dy, dz, *_dims(Q), _stream()))


def initial_bcs(Q):
    L = _lib.load()
    check(L.imhd_initial_bcs(_dev(Q), *_dims(Q), _stream()))


def init_grids(bounds, Nx, Ny, Nz, device="cuda"):
    import torch

    L = _lib.load()
    x, y, z = (torch.empty(n, dtype=torch.float32, device=device) for n in (Nx, Ny, Nz))
    check(L.imhd_init_grids(_dev(x), _dev(y), _dev(z), *[float(b) for b in bounds], Nx, Ny, Nz, _stream()))
    return x, y, z


def init_screwpinch_stride(J0, x, y, z):
    import torch

    L = _lib.load()
    Q = torch.empty((8, len(z), len(x), len(y)), dtype=torch.float32, device=x.device)
    check(L.imhd_init_screwpinch_stride(_dev(Q), J0, _dev(x), _dev(y), _dev(z), len(x), len(y), len(z), _stream()))
    return Q


def init_cubic_bennett_vortex_m0(k, A, x, y, z):
    import torch

    L = _lib.load()
    Q = torch.empty((8, len(z), len(x), len(y)), dtype=torch.float32, device=x.device)
    check(L.imhd_init_cubic_bennett_vortex_m0(_dev(Q), k, A, _dev(x), _dev(y), _dev(z), len(x), len(y), len(z),
                                              _stream()))
    return Q


def _init_ic(fn_name, x, y, z, *scalars, prefill=None):
    import torch

    L = _lib.load()
    shape = (8, len(z), len(x), len(y))
    Q = torch.empty(shape, dtype=torch.float32, device=x.device) if prefill is None else prefill.clone()
    check(getattr(L, fn_name)(_dev(Q), *scalars, _dev(x), _dev(y), _dev(z), len(x), len(y), len(z), _stream()))
    return Q


def init_cubic_bennett_vortex(x, y, z):
    """CubicBennettVortex (lib/on-device/initialize_od.cu:59-130)."""
    return _init_ic("imhd_init_cubic_bennett_vortex", x, y, z)


def init_zpinch(r_max_coeff, x, y, z):
    """ZPinch (lib/on-device/initialize_od.cu:347-424)."""
    return _init_ic("imhd_init_zpinch", x, y, z, r_max_coeff)


def init_screwpinch(J0, r_max_coeff, x, y, z, prefill=None):
    """ScrewPinch (lib/on-device/initialize_od.cu:207-267).  Outside the pinch only rho is written; the other seven
    variables keep the contents of ``prefill`` (the reference leaves its cudaMalloc as it found it)."""
    import torch

    if prefill is None:
        prefill = torch.zeros((8, len(z), len(x), len(y)), dtype=torch.float32, device=x.device)
    return _init_ic("imhd_init_screwpinch", x, y, z, J0, r_max_coeff, prefill=prefill)


# ---- registry (include/on-device/utils/configurers.hpp) ------------------------------------------------
REG_INITIALIZER, REG_CORRECTOR, REG_PREDICTOR, REG_FLUID_BCS, REG_PREDICTOR_BCS = range(5)


def registry_names(kind: int) -> list[str]:
    L = _lib.load()
    return [L.imhd_registry_name(kind, q).decode() for q in range(L.imhd_registry_count(kind))]


def registry_resolve_path(corrector: str, predictor: str, fluid_bcs: str = "pcrw-xy_pbc-z",
                          predictor_bcs: str = "pbc-z") -> int:
    L = _lib.load()
    path = C.c_int(-1)
    check(L.imhd_registry_resolve_path(corrector.encode(), predictor.encode(), fluid_bcs.encode(),
                                       predictor_bcs.encode(), C.byref(path)))
    return path.value


# ---- CFL / stability scan (src/on-device/utils/compute_stability.cpp) -----------------------------------
def stability_scan(Q, slab: Slab) -> dict:
    """Scan the owned planes of a slab state on its device; returns max_lhs, argmax (i, j, k) in global indices,
    violations and the reference's dt proposal.  Synchronises."""
    out = _lib.Stability()
    check(_lib.load().imhd_stability_scan(_dev(Q), C.byref(slab), C.byref(out), _stream()))
    return {"max_lhs": out.max_lhs, "argmax_ijk": (out.i, out.j, out.k), "violations": int(out.violations),
            "dt_new": out.dt_new}


# ---- fused step -----------------------------------------------------------------------------------
def make_slab(Nx, Ny, Nz, path, D, dt, dx, dy, dz, k0=0, nzl=None, ghosts=0, corner_e=0.0) -> Slab:
    return Slab(Nx, Ny, Nz, k0, Nz if nzl is None else nzl, ghosts, path, D, dt, dx, dy, dz, corner_e)


def qint_plane(Q, k, slab: Slab, out=None):
    """Predictor plane Qint(.,.,k) (global k) from a slab array -> (8, Nx, Ny)."""
    import torch

    L = _lib.load()
    if out is None:
        out = torch.empty((8, slab.Nx, slab.Ny), dtype=torch.float32, device=Q.device)
    check(L.imhd_qint_plane(_dev(Q), _dev(out, (8, slab.Nx, slab.Ny)), k, C.byref(slab), _stream()))
    return out


def step_fused(Qin, Qout, qint_lo, qint_hi, qint_wrap, slab: Slab):
    L = _lib.load()
    wrap = _dev(qint_wrap) if qint_wrap is not None else None
    check(L.imhd_step_fused(_dev(Qin), _dev(Qout, Qin.shape), _dev(qint_lo), _dev(qint_hi), wrap, C.byref(slab),
                            _stream()))


def step_fused_planes(Qin, Qout, qint_lo, qint_hi, qint_wrap, slab: Slab, kfrom: int, kto: int):
    """Output planes [kfrom, kto) of the slab only (same bits as the full step for any split)."""
    L = _lib.load()
    wrap = _dev(qint_wrap) if qint_wrap is not None else None
    check(L.imhd_step_fused_planes(_dev(Qin), _dev(Qout, Qin.shape), _dev(qint_lo), _dev(qint_hi), wrap, C.byref(slab),
                                   kfrom, kto, _stream()))


def step_fused_ends(Qin, Qout, qint_lo, qint_hi, qint_wrap, slab: Slab, kfrom: int, kmid1: int, kmid2: int, kto: int):
    """Output planes [kfrom, kmid1) and [kmid2, kto) of the slab in one launch (both slab ends of the multi-GPU loop)."""
    L = _lib.load()
    wrap = _dev(qint_wrap) if qint_wrap is not None else None
    check(L.imhd_step_fused_ends(_dev(Qin), _dev(Qout, Qin.shape), _dev(qint_lo), _dev(qint_hi), wrap, C.byref(slab),
                                 kfrom, kmid1, kmid2, kto, _stream()))


def wall_energy_fixed_point(e: float, max_iter: int) -> float:
    return float(_lib.load().imhd_wall_energy_fixed_point(e, max_iter))


def step_full_domain(Qin, Qout, path, D, dt, dx, dy, dz, corner_e=0.0):
    """One fused step of a whole-domain (8,Nz,Nx,Ny) array on one GPU, including the periodic wrap planes."""
    Nx, Ny, Nz = _dims(Qin)
    s = make_slab(Nx, Ny, Nz, path, D, dt, dx, dy, dz, corner_e=corner_e)
    q0 = qint_plane(Qin, 0, s)  # Qint(0) == Qint(Nz-1): both ends of the slab
    wrap = qint_plane(Qin, Nz - 2, s) if path == PATH_B else None
    step_fused(Qin, Qout, q0, q0, wrap, s)


def nccl_unique_id() -> bytes:
    """128 bytes identifying one NCCL clique: made by ONE process, handed to every process of Context.slab."""
    buf = C.create_string_buffer(128)
    check(_lib.load().imhd_nccl_unique_id(buf, 128))
    return buf.raw


def launch_count() -> int:
    return int(_lib.load().imhd_launch_count())


# ---- context ------------------------------------------------------------------------------------------
class Context:
    """``imhd_ctx``: what src/on-device/main.cu / no_diffusion.cu own between their IC and their time loop."""

    def __init__(self, Nx: int, Ny: int, Nz: int, device: int = 0, _handle=None):
        self.L = _lib.load()
        self.Nx, self.Ny, self.Nz = Nx, Ny, Nz
        self.h = _handle if _handle is not None else self.L.imhd_create(Nx, Ny, Nz, device)
        if not self.h:
            raise _lib.ImhdError(-1, self.L.imhd_last_error().decode())

    # ---- multi-GPU contexts: the z-slab time loop in C++ behind the same calls (imhd_create_multi / imhd_create_slab) ----
    @classmethod
    def multi(cls, Nx: int, Ny: int, Nz: int, n_gpus: int, devices=None):
        """The whole domain as n_gpus z-slabs driven from this process (one slab per device)."""
        L = _lib.load()
        arr = (C.c_int * n_gpus)(*(devices if devices is not None else range(n_gpus)))
        return cls(Nx, Ny, Nz, _handle=L.imhd_create_multi(Nx, Ny, Nz, n_gpus, arr) or 0)

    @classmethod
    def slab(cls, Nx: int, Ny: int, Nz: int, rank: int, world: int, device: int, unique_id: bytes):
        """Slab `rank` of `world` in this process (one process per GPU); unique_id from nccl_unique_id() of one process."""
        L = _lib.load()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        return cls(Nx, Ny, Nz, _handle=L.imhd_create_slab(Nx, Ny, Nz, rank, world, device, buf) or 0)

    @property
    def num_slabs(self) -> int:
        return int(self.L.imhd_ctx_num_slabs(self.h))

    def slab_extent(self, q: int = 0):
        """(k0, nzl, device) of local slab q."""
        k0, nzl, dev = C.c_int(), C.c_int(), C.c_int()
        check(self.L.imhd_ctx_slab_extent(self.h, q, C.byref(k0), C.byref(nzl), C.byref(dev)))
        return k0.value, nzl.value, dev.value

    def set_state_local(self, Q: np.ndarray, q: int = 0):
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        assert Q.shape == (8, self.slab_extent(q)[1], self.Nx, self.Ny)
        check(self.L.imhd_ctx_set_state_local(self.h, q, Q.ctypes.data_as(C.c_void_p)))
        check(self.L.imhd_ctx_synchronize(self.h))

    def get_state_local(self, q: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((8, self.slab_extent(q)[1], self.Nx, self.Ny), np.float32)
        check(self.L.imhd_ctx_get_state_local(self.h, q, out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.imhd_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def shape(self):
        return (8, self.Nz, self.Nx, self.Ny)

    def init_grids(self, x_min, x_max, y_min, y_max, z_min, z_max):
        check(self.L.imhd_ctx_init_grids(self.h, x_min, x_max, y_min, y_max, z_min, z_max))

    def init_screwpinch_stride(self, J0):
        check(self.L.imhd_ctx_init_screwpinch_stride(self.h, J0))

    def init_cubic_bennett_vortex_m0(self, k, A):
        check(self.L.imhd_ctx_init_cubic_bennett_vortex_m0(self.h, k, A))

    def initialize(self, sim_type: str, *params: float):
        """SimulationInitializer::initialize (configurers.hpp:32-39): IC kernel by registry key."""
        arr = (C.c_float * max(1, len(params)))(*params)
        check(self.L.imhd_ctx_initialize(self.h, sim_type.encode(), arr, len(params)))

    def stability(self, dt: float) -> dict:
        out = _lib.Stability()
        check(self.L.imhd_ctx_stability(self.h, dt, C.byref(out)))
        return {"max_lhs": out.max_lhs, "argmax_ijk": (out.i, out.j, out.k), "violations": int(out.violations),
                "dt_new": out.dt_new}

    def step_adaptive(self, nsteps: int, every: int = 10, cfl_target: float = 0.5, dt_max: float = 1e30) -> np.ndarray:
        """nsteps fused steps with the time step following the CFL scan (imhd_ctx_step_adaptive: scanned every `every`
        steps without stalling the loop, applied one group of steps later).  Returns the dt of every step."""
        used = np.zeros(nsteps, np.float32)
        check(self.L.imhd_ctx_step_adaptive(self.h, nsteps, every, cfl_target, dt_max, used.ctypes.data_as(C.c_void_p)))
        return used

    def set_dt(self, dt: float):
        check(self.L.imhd_ctx_set_dt(self.h, dt))

    def set_state(self, Q: np.ndarray):
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        assert Q.shape == self.shape
        check(self.L.imhd_ctx_set_state(self.h, Q.ctypes.data_as(C.c_void_p)))
        check(self.L.imhd_ctx_synchronize(self.h))  # Q may be a temporary

    def set_spacing(self, dx, dy, dz):
        check(self.L.imhd_ctx_set_spacing(self.h, dx, dy, dz))

    def prime(self, path, D, dt):
        check(self.L.imhd_ctx_prime(self.h, path, D, dt))

    def step(self, nsteps=1):
        check(self.L.imhd_ctx_step(self.h, nsteps))

    def step_granular(self, nsteps=1):
        check(self.L.imhd_ctx_step_granular(self.h, nsteps))

    def synchronize(self):
        check(self.L.imhd_ctx_synchronize(self.h))

    def get_state(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, np.float32)
        check(self.L.imhd_ctx_get_state(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_grids(self):
        x, y, z = (np.empty(n, np.float32) for n in (self.Nx, self.Ny, self.Nz))
        check(self.L.imhd_ctx_get_grids(self.h, *(a.ctypes.data_as(C.c_void_p) for a in (x, y, z))))
        return x, y, z

    # ---- output (src/on-device/utils/phdf5_write_all.cpp, hdf5_write_grid.cpp; asynchronous, see imhd_ctx_write_frame) ----
    def write_frame(self, directory: str, frame: int):
        """Queue the current state as <directory>/fluidvars_<frame>.h5; returns at once."""
        d = directory if directory.endswith("/") else directory + "/"
        check(self.L.imhd_ctx_write_frame(self.h, d.encode(), frame))

    def write_grid(self, directory: str):
        d = directory if directory.endswith("/") else directory + "/"
        check(self.L.imhd_ctx_write_grid(self.h, d.encode()))

    def flush_output(self):
        """Wait until every queued frame is on disk; raises if a write failed."""
        check(self.L.imhd_ctx_flush_output(self.h))

    def run_host(self, Q_in: np.ndarray, Q_out: np.ndarray, path, D, dt, dx, dy, dz, nsteps):
        check(self.L.imhd_run_host(self.h, Q_in.ctypes.data_as(C.c_void_p), Q_out.ctypes.data_as(C.c_void_p), path, D,
                                   dt, dx, dy, dz, nsteps))
