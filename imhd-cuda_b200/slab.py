"""z-slab domain decomposition: one slab per GPU, one process per GPU (torch.distributed / NCCL).

The reference is single-GPU; the decomposition is new work.  k is the slowest index of
the reference layout, so a slab is one contiguous plane range per variable.  Rank r owns the global
planes [k0, k1) and stores them with one ghost plane on each side as an (8, nzl+2, Nx, Ny) array.

Per time step and rank (ring neighbours, periodic in z):
  1. compute the two predictor planes the neighbours need (``imhd_qint_plane``) and exchange them:
       up-going   Qint(k1-1)   -> rank r+1's ``qint_lo``   (the last rank sends Qint(Nz-2) == Qint(-1), which
                                  rank 0 uses as ``qint_wrap`` for the k=0 face; rank 0's ``qint_lo`` is its own Qint(0))
       down-going Qint(k0)     -> rank r-1's ``qint_hi``   (rank 0 sends Qint(0): Qint(Nz-1) == Qint(0))
  2. one fused kernel over the owned planes (``imhd_step_fused``)
  3. exchange the new boundary planes of Q into the neighbours' ghost planes; for path A the plane the
     last rank sends up is the periodic copy Q[.,.,0] <- Q[.,.,Nz-1] (lib/on-device/kernels_fluidbcs.cu:498-510)
     and lands in rank 0's OWNED plane 0.
No reduction is needed anywhere (the reference has no global dt control).  Results are bit-identical for
every number of ranks: each value is computed by the same device function from the same inputs.

The communication layer is injected (``comm``) so the host logic runs under gloo on CPU in the tests.
"""
from __future__ import annotations

from dataclasses import dataclass

from ._lib import PATH_A, PATH_B


@dataclass(frozen=True)
class SlabLayout:
    """Which global planes rank ``rank`` of ``world`` owns, and who its ring neighbours are."""

    Nz: int
    world: int
    rank: int

    def __post_init__(self):
        if self.world < 1 or not (0 <= self.rank < self.world):
            raise ValueError("bad rank/world")
        if self.Nz // self.world < 3:
            raise ValueError(f"Nz={self.Nz} is too thin for {self.world} slabs (need >= 3 planes per slab)")

    @property
    def k0(self) -> int:
        return (self.Nz * self.rank) // self.world

    @property
    def k1(self) -> int:
        return (self.Nz * (self.rank + 1)) // self.world

    @property
    def nzl(self) -> int:
        return self.k1 - self.k0

    @property
    def up(self) -> int:
        return (self.rank + 1) % self.world

    @property
    def down(self) -> int:
        return (self.rank - 1) % self.world

    @property
    def up_plane(self) -> int:
        """Global index of the predictor plane sent to the rank above (its qint_lo)."""
        return self.k1 - 1 if self.rank < self.world - 1 else self.Nz - 2

    @property
    def down_plane(self) -> int:
        """Global index of the predictor plane sent to the rank below (its qint_hi)."""
        return self.k0

    def local(self, k: int) -> int:
        """Array plane of global plane k (ghosted array: plane 0 is global k0-1)."""
        return k - self.k0 + 1


class TorchComm:
    """Ring exchange over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def ring_exchange(self, send_up, send_down, recv_from_down, recv_from_up, up: int, down: int):
        """send_up -> rank `up`, send_down -> rank `down`; recv_from_down <- `down`, recv_from_up <- `up`."""
        d = self.dist
        if self.world == 2:
            # up == down: order the two messages between the same pair by tag-free pairing
            ops = [d.P2POp(d.isend, send_up, up, self.group), d.P2POp(d.irecv, recv_from_down, down, self.group),
                   d.P2POp(d.isend, send_down, down, self.group), d.P2POp(d.irecv, recv_from_up, up, self.group)]
        else:
            ops = [d.P2POp(d.isend, send_up, up, self.group), d.P2POp(d.isend, send_down, down, self.group),
                   d.P2POp(d.irecv, recv_from_down, down, self.group), d.P2POp(d.irecv, recv_from_up, up, self.group)]
        for req in d.batch_isend_irecv(ops):
            req.wait()


class SlabSolver:
    """Time loop of one slab.  ``compute`` supplies qint_plane / step_fused (the C ABI through ops.py on a GPU)."""

    def __init__(self, Nx, Ny, Nz, path, D, dt, dx, dy, dz, comm=None, compute=None, device="cuda", corner_e=0.0):
        import torch

        self.torch = torch
        self.comm = comm
        world = comm.world if comm is not None else 1
        rank = comm.rank if comm is not None else 0
        self.layout = SlabLayout(Nz, world, rank)
        self.Nx, self.Ny, self.Nz, self.path = Nx, Ny, Nz, path
        self.params = (D, dt, dx, dy, dz)
        self.corner_e = corner_e
        if compute is None:
            from . import ops as compute
        self.compute = compute
        L = self.layout
        shape = (8, L.nzl + 2, Nx, Ny)
        self.Q = [torch.zeros(shape, dtype=torch.float32, device=device) for _ in range(2)]
        self.cur = 0
        pl = (8, Nx, Ny)
        self.qint_lo, self.qint_hi, self.send_up, self.send_down, self.recv_lo, self.recv_hi = (
            torch.zeros(pl, dtype=torch.float32, device=device) for _ in range(6))
        self.slab = compute.make_slab(Nx, Ny, Nz, path, D, dt, dx, dy, dz, k0=L.k0, nzl=L.nzl, ghosts=1,
                                      corner_e=corner_e)

    # ---- state in / out ----------------------------------------------------------------------------
    @property
    def state(self):
        """Owned planes of the current state, (8, nzl, Nx, Ny) view."""
        return self.Q[self.cur][:, 1:-1]

    def load_global(self, Qglobal):
        """Fill owned + ghost planes from a full (8,Nz,Nx,Ny) host/device array (every rank holds a copy)."""
        L, t = self.layout, self.torch
        src = t.as_tensor(Qglobal)
        lo, hi = max(L.k0 - 1, 0), min(L.k1 + 1, self.Nz)
        self.Q[self.cur][:, lo - L.k0 + 1: hi - L.k0 + 1].copy_(src[:, lo:hi])

    # ---- exchanges -----------------------------------------------------------------------------------
    def exchange_qint(self):
        L, c = self.layout, self.compute
        Q = self.Q[self.cur]
        if self.comm is None or L.world == 1:
            c.qint_plane(Q, 0, self.slab, out=self.qint_hi)          # Qint(0) == Qint(Nz-1)
            self.lo, self.hi, self.wrap = self.qint_hi, self.qint_hi, None
            if self.path == PATH_B:
                c.qint_plane(Q, self.Nz - 2, self.slab, out=self.qint_lo)  # Qint(Nz-2) == Qint(-1)
                self.wrap = self.qint_lo
            return
        c.qint_plane(Q, L.up_plane, self.slab, out=self.send_up)
        c.qint_plane(Q, L.down_plane, self.slab, out=self.send_down)
        self.comm.ring_exchange(self.send_up, self.send_down, self.qint_lo, self.qint_hi, L.up, L.down)
        if L.rank == 0:  # received Qint(Nz-2) from the last rank; the plane below plane 1 is this rank's own Qint(0)
            self.lo, self.hi, self.wrap = self.send_down, self.qint_hi, self.qint_lo
        else:
            self.lo, self.hi, self.wrap = self.qint_lo, self.qint_hi, None

    def exchange_ghosts(self):
        """New boundary planes of Q -> neighbours' ghost planes (+ the path A periodic copy)."""
        L = self.layout
        Q = self.Q[self.cur]
        if self.comm is None or L.world == 1:
            return  # the fused kernel wrote plane 0 itself (path A) and nothing reads a ghost plane
        self.send_up.copy_(Q[:, L.nzl])   # owned top plane k1-1
        self.send_down.copy_(Q[:, 1])     # owned bottom plane k0
        self.comm.ring_exchange(self.send_up, self.send_down, self.recv_lo, self.recv_hi, L.up, L.down)
        if L.rank > 0:
            Q[:, 0].copy_(self.recv_lo)
        elif self.path == PATH_A:
            Q[:, 1].copy_(self.recv_lo)   # PBCs: global plane 0 <- global plane Nz-1
        if L.rank < L.world - 1:
            Q[:, L.nzl + 1].copy_(self.recv_hi)

    # ---- time loop --------------------------------------------------------------------------------------
    def step(self, nsteps=1):
        for _ in range(nsteps):
            self.exchange_qint()
            Qin, Qout = self.Q[self.cur], self.Q[1 - self.cur]
            self.compute.step_fused(Qin, Qout, self.lo, self.hi, self.wrap, self.slab)
            self.cur = 1 - self.cur
            self.exchange_ghosts()


__all__ = ["SlabLayout", "SlabSolver", "TorchComm", "PATH_A", "PATH_B"]
