"""z-slab domain decomposition: one slab per GPU, one process per GPU (torch.distributed / NCCL).

The reference is single-GPU; the decomposition is new work.  k is the slowest index of
the reference layout, so a slab is one contiguous plane range per variable.  Rank r owns the global
planes [k0, k1) and stores them with one ghost plane on each side as an (8, nzl+2, Nx, Ny) array.

Per time step and rank (ring neighbours, periodic in z):
  1. compute the two predictor planes the neighbours need (``imhd_qint_plane``) and exchange them:
       up-going   Qint(k1-1)   -> rank r+1's ``qint_lo``   (the last rank sends Qint(Nz-2) == Qint(-1), which
                                  rank 0 uses as ``qint_wrap`` for the k=0 face; rank 0's ``qint_lo`` is its own Qint(0))
       down-going Qint(k0)     -> rank r-1's ``qint_hi``   (rank 0 sends Qint(0): Qint(Nz-1) == Qint(0))
  2. the fused kernel over the owned planes (``imhd_step_fused_planes``: slab ends first, then the interior)
  3. exchange the new boundary planes of Q into the neighbours' ghost planes; for path A the plane the
     last rank sends up is the periodic copy Q[.,.,0] <- Q[.,.,Nz-1] (lib/on-device/kernels_fluidbcs.cu:498-510)
     and lands in rank 0's OWNED plane 0.
Steps 1 and 3 run on a side stream underneath the interior launch of step 2 (the predictor planes of the next
step only need the new end planes), so the interior never waits for a message.
The time step needs no reduction (the reference has no global dt control).  Results are bit-identical for
every number of ranks: each value is computed by the same device function from the same inputs.
The optional CFL scan (``SlabSolver.stability``, replacing the reference's forked host scanner) is the one place
with a reduction: max of the per-slab maxima, sum of the violation counts, one small all-gather.

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

    def allgather_row(self, row):
        """Every rank contributes one row of numbers; returns the rows of all ranks in rank order (float64)."""
        import torch

        dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
        mine = torch.tensor(row, dtype=torch.float64, device=dev)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(out, mine, group=self.group)
        return [o.tolist() for o in out]


def combine_stability(rows, dt, alpha=0.1):
    """Per-slab scan results [max_lhs, i, j, k, violations] in rank (= ascending k) order -> the scan of the whole
    domain: the largest LHS (the first one in the reference's scan order k, i, j on ties), the total number of
    violations and dt_new = alpha * dt / max (src/on-device/utils/compute_stability.cpp:139-141)."""
    best = None
    for r in rows:
        if best is None or r[0] > best[0]:
            best = r
    mx = float(best[0])
    return {"max_lhs": mx, "argmax_ijk": (int(best[1]), int(best[2]), int(best[3])),
            "violations": int(sum(int(r[4]) for r in rows)), "dt_new": alpha * dt / mx if mx > 0 else 0.0}


class SlabSolver:
    """Time loop of one slab.  ``compute`` supplies qint_plane / step_fused_planes (the C ABI through ops.py on a GPU).

    With more than one rank on GPUs the step is software-pipelined: the planes next to the slab ends are computed
    first, then -- on a side stream, under the interior launch -- their ghost exchange, the two predictor planes of
    the NEW state and the exchange of those.  The interior never waits for a message."""

    EDGE = 8  # planes computed ahead at each slab end (>= 3: the predictor planes of the new state need them)

    def __init__(self, Nx, Ny, Nz, path, D, dt, dx, dy, dz, comm=None, compute=None, device="cuda", corner_e=None):
        import torch

        self.torch = torch
        self.comm = comm
        world = comm.world if comm is not None else 1
        rank = comm.rank if comm is not None else 0
        self.layout = SlabLayout(Nz, world, rank)
        self.Nx, self.Ny, self.Nz, self.path = Nx, Ny, Nz, path
        self.params = (D, dt, dx, dy, dz)
        # path B: the wall value the column (Nx-1,Ny-1) carries at k = 0 and k = Nz-1 from step 1 on (B-8).  None = derive
        # it from the loaded state (load_global / derive_corner_e) instead of trusting a caller-supplied number.
        self._corner_given = corner_e is not None
        self.corner_e = 0.0 if corner_e is None else corner_e
        if compute is None:
            from . import ops as compute
        self.compute = compute
        L = self.layout
        shape = (8, L.nzl + 2, Nx, Ny)
        self.Q = [torch.zeros(shape, dtype=torch.float32, device=device) for _ in range(2)]
        self.cur = 0
        pl = (8, Nx, Ny)
        mk = lambda: torch.zeros(pl, dtype=torch.float32, device=device)  # noqa: E731
        # predictor planes are double buffered: the set for step n+1 is produced while step n still reads its own
        self.qsets = [{"lo": mk(), "hi": mk(), "up": mk(), "down": mk()} for _ in range(2)]
        self.qcur, self.q_ready = 0, False
        self.send_up, self.send_down, self.recv_lo, self.recv_hi = mk(), mk(), mk(), mk()
        self.slab = compute.make_slab(Nx, Ny, Nz, path, D, dt, dx, dy, dz, k0=L.k0, nzl=L.nzl, ghosts=1,
                                      corner_e=corner_e)
        self.overlap = world > 1 and str(device).startswith("cuda") and L.nzl >= 2 * self.EDGE + 2
        if self.overlap:
            self.side = torch.cuda.Stream()
            self.ev_edges, self.ev_comm = torch.cuda.Event(), torch.cuda.Event()

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
        self.q_ready = False
        if not self._corner_given and self.path == PATH_B:
            self.set_corner_e(float(src[7, 0, self.Nx - 1, self.Ny - 1]))

    def set_corner_e(self, e0: float):
        """corner_e from e(Nx-1, Ny-1, k=0) of the INITIAL state: the fixed point of the reference's wall-energy map
        (kernels_fluidbcs.cu:173).  With several ranks the owner of plane 0 (rank 0) supplies e0 to all."""
        self.corner_e = float(self.compute.wall_energy_fixed_point(e0, self.Nx)) if hasattr(self.compute, "wall_energy_fixed_point") else 0.0
        L = self.layout
        D, dt, dx, dy, dz = self.params
        self.slab = self.compute.make_slab(self.Nx, self.Ny, self.Nz, self.path, D, dt, dx, dy, dz, k0=L.k0, nzl=L.nzl, ghosts=1,
                                           corner_e=self.corner_e)

    def derive_corner_e(self):
        """After the slabs were filled on the device (not through load_global): rank 0 reads e(Nx-1,Ny-1,0) and every
        rank receives it."""
        if self.path != PATH_B:
            return
        L = self.layout
        e0 = float(self.Q[self.cur][7, 1, self.Nx - 1, self.Ny - 1]) if L.rank == 0 else 0.0
        if self.comm is not None and L.world > 1:
            e0 = self.comm.allgather_row([e0])[0][0]
        self.set_corner_e(e0)

    # ---- exchanges -----------------------------------------------------------------------------------
    def exchange_qint(self, Q, qs):
        """Fill the predictor-plane set `qs` for a step that starts from state array Q; returns (lo, hi, wrap)."""
        L, c = self.layout, self.compute
        if self.comm is None or L.world == 1:
            c.qint_plane(Q, 0, self.slab, out=qs["hi"])          # Qint(0) == Qint(Nz-1)
            wrap = None
            if self.path == PATH_B:
                c.qint_plane(Q, self.Nz - 2, self.slab, out=qs["lo"])  # Qint(Nz-2) == Qint(-1)
                wrap = qs["lo"]
            return qs["hi"], qs["hi"], wrap
        c.qint_plane(Q, L.up_plane, self.slab, out=qs["up"])
        c.qint_plane(Q, L.down_plane, self.slab, out=qs["down"])
        self.comm.ring_exchange(qs["up"], qs["down"], qs["lo"], qs["hi"], L.up, L.down)
        if L.rank == 0:  # received Qint(Nz-2) from the last rank; the plane below plane 1 is this rank's own Qint(0)
            return qs["down"], qs["hi"], qs["lo"]
        return qs["lo"], qs["hi"], None

    def exchange_ghosts(self, Q):
        """New boundary planes of Q -> neighbours' ghost planes (+ the path A periodic copy)."""
        L = self.layout
        if self.comm is None or L.world == 1:
            return  # the fused step wrote plane 0 itself (path A) and nothing reads a ghost plane
        self.send_up.copy_(Q[:, L.nzl])   # owned top plane k1-1
        self.send_down.copy_(Q[:, 1])     # owned bottom plane k0
        self.comm.ring_exchange(self.send_up, self.send_down, self.recv_lo, self.recv_hi, L.up, L.down)
        if L.rank > 0:
            Q[:, 0].copy_(self.recv_lo)
        elif self.path == PATH_A:
            Q[:, 1].copy_(self.recv_lo)   # PBCs: global plane 0 <- global plane Nz-1
        if L.rank < L.world - 1:
            Q[:, L.nzl + 1].copy_(self.recv_hi)

    # ---- CFL scan ---------------------------------------------------------------------------------------
    def stability(self, dt=None):
        """Stability criterion of the current state over the WHOLE domain (every rank returns the same dict)."""
        D, dt0, dx, dy, dz = self.params
        dt = dt0 if dt is None else dt
        L, c = self.layout, self.compute
        slab = self.slab if dt == dt0 else c.make_slab(self.Nx, self.Ny, self.Nz, self.path, D, dt, dx, dy, dz, k0=L.k0,
                                                       nzl=L.nzl, ghosts=1, corner_e=self.corner_e)
        r = c.stability_scan(self.Q[self.cur], slab)
        row = [r["max_lhs"], *r["argmax_ijk"], r["violations"]]
        rows = [row] if self.comm is None or L.world == 1 else self.comm.allgather_row(row)
        return combine_stability(rows, dt)

    # ---- time loop --------------------------------------------------------------------------------------
    def step(self, nsteps=1):
        L, c, t = self.layout, self.compute, self.torch
        for _ in range(nsteps):
            Qin, Qout = self.Q[self.cur], self.Q[1 - self.cur]
            if not self.q_ready:
                self.planes = self.exchange_qint(Qin, self.qsets[self.qcur])
                self.q_ready = True
            lo, hi, wrap = self.planes
            if not self.overlap:
                c.step_fused_planes(Qin, Qout, lo, hi, wrap, self.slab, L.k0, L.k1)
                self.cur = 1 - self.cur
                self.exchange_ghosts(Qout)
                self.q_ready = False
                continue
            E = self.EDGE
            main = t.cuda.current_stream()
            c.step_fused_planes(Qin, Qout, lo, hi, wrap, self.slab, L.k0, L.k0 + E)
            c.step_fused_planes(Qin, Qout, lo, hi, wrap, self.slab, L.k1 - E, L.k1)
            self.ev_edges.record(main)
            c.step_fused_planes(Qin, Qout, lo, hi, wrap, self.slab, L.k0 + E, L.k1 - E)
            with t.cuda.stream(self.side):   # under the interior launch: halo of the new state, predictor planes of the next step
                self.side.wait_event(self.ev_edges)
                self.exchange_ghosts(Qout)
                self.qcur = 1 - self.qcur
                self.planes = self.exchange_qint(Qout, self.qsets[self.qcur])
                self.ev_comm.record(self.side)
            main.wait_event(self.ev_comm)
            self.cur = 1 - self.cur


__all__ = ["SlabLayout", "SlabSolver", "TorchComm", "combine_stability", "PATH_A", "PATH_B"]
