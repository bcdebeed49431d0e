"""Host-side logic of the z-slab decomposition under torch.distributed (gloo, CPU, world_size 2 and 3):
slab layout, the ring exchange of predictor planes and ghost planes including the periodic wrap, and the
order in which SlabSolver.step drives compute and communication.  The compute back end is a recording stub
(the real one is the C ABI and needs a GPU; the bit-identity of decomposed runs is a -m gpu test)."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layout_partitions_the_domain():
    slab = importlib.import_module("imhd-cuda_b200.slab")
    for Nz, world in ((592, 1), (592, 8), (128, 3), (4736, 8), (26, 3)):
        ls = [slab.SlabLayout(Nz, world, r) for r in range(world)]
        assert ls[0].k0 == 0 and ls[-1].k1 == Nz
        assert all(a.k1 == b.k0 for a, b in zip(ls, ls[1:]))
        assert all(l.nzl >= 3 for l in ls)
        assert ls[-1].up == 0 and ls[0].down == world - 1
        assert ls[-1].up_plane == Nz - 2 and ls[0].down_plane == 0
        for l in ls[:-1]:
            assert l.up_plane == l.k1 - 1
    with pytest.raises(ValueError):
        slab.SlabLayout(8, 4, 0)


class StubCompute:
    """qint_plane(k) fills the plane with 1000+k; step_fused writes (old value + 1) on owned planes and records
    which predictor planes it was handed."""

    PATH_B = 1

    def __init__(self):
        self.seen = []

    class _Slab:
        pass

    def make_slab(self, Nx, Ny, Nz, path, D, dt, dx, dy, dz, k0=0, nzl=None, ghosts=0, corner_e=0.0):
        s = self._Slab()
        s.Nx, s.Ny, s.Nz, s.k0, s.nzl, s.path = Nx, Ny, Nz, k0, nzl, path
        return s

    def qint_plane(self, Q, k, slab, out=None):
        out.fill_(1000.0 + k)
        return out

    def stability_scan(self, Q, slab):
        # rank-dependent fake scan: the maximum sits in the LAST slab, ties with the one before it
        top = 0.25 * min(slab.k0 // slab.nzl + 1, slab.Nz // slab.nzl - 1) if slab.Nz // slab.nzl > 2 else 0.25 * (slab.k0 // slab.nzl + 1)
        return {"max_lhs": top, "argmax_ijk": (1, 2, slab.k0 + 1), "violations": slab.k0 + 3, "dt_new": 0.0}

    def step_fused_planes(self, Qin, Qout, lo, hi, wrap, slab, kfrom, kto):
        assert (kfrom, kto) == (slab.k0, slab.k0 + slab.nzl)  # CPU path: no overlap, one launch over the owned planes
        self.seen.append((float(lo[0, 0, 0]), float(hi[0, 0, 0]), None if wrap is None else float(wrap[0, 0, 0])))
        Qout[:, 1:-1] = Qin[:, 1:-1] + 1.0
        if slab.path == 0 and slab.k0 == 0:
            Qout[:, 1] = -777.0  # path A never computes plane 0; the exchange must fill it


def _worker(rank, world, path, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slab = importlib.import_module("imhd-cuda_b200.slab")
        Nx, Ny, Nz = 4, 4, 6 * world
        comp = StubCompute()
        s = slab.SlabSolver(Nx, Ny, Nz, path, 0.0, 1e-4, 0.1, 0.1, 0.1, comm=slab.TorchComm(), compute=comp, device="cpu")
        L = s.layout
        glob = torch.arange(Nz, dtype=torch.float32).view(1, Nz, 1, 1).expand(8, Nz, Nx, Ny).contiguous()  # Q(k) = k
        s.load_global(glob)
        s.step(2)
        Q = s.Q[s.cur]
        stab = s.stability()
        res = {"rank": rank, "stab": stab, "seen": comp.seen, "owned": Q[0, 1:-1, 0, 0].tolist(), "lo_ghost": float(Q[0, 0, 0, 0]),
               "hi_ghost": float(Q[0, -1, 0, 0]), "k0": L.k0, "k1": L.k1}
        q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("path", [0, 1])
def test_ring_exchange_under_gloo(world, path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world * 10 + path + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, path, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r["rank"])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    Nz = 6 * world
    # CFL scan: every rank holds the same combined result; on a tie the first slab in scan order (smallest k) wins
    tops = [0.25 * (min(rk + 1, world - 1) if world > 2 else rk + 1) for rk in range(world)]
    first = tops.index(max(tops))
    want = {"max_lhs": max(tops), "argmax_ijk": (1, 2, 6 * first + 1), "violations": sum(6 * rk + 3 for rk in range(world)),
            "dt_new": 0.1 * 1e-4 / max(tops)}
    for r in res:
        assert r["stab"] == pytest.approx(want), r["stab"]
    for r in res:
        k0, k1, rank = r["k0"], r["k1"], r["rank"]
        # predictor planes handed to the fused step: lo = Qint(k0-1) (own Qint(0) on rank 0), hi = Qint(k1) (Qint(0) on the last rank),
        # wrap = Qint(Nz-2) on rank 0 only
        exp_lo = 1000.0 + (0 if rank == 0 else k0 - 1)
        exp_hi = 1000.0 + (0 if rank == world - 1 else k1)
        exp_wrap = 1000.0 + Nz - 2 if rank == 0 else None
        assert r["seen"] == [(exp_lo, exp_hi, exp_wrap)] * 2, r
        # after 2 steps every owned plane k holds k+2, except path A's plane 0 = copy of plane Nz-1
        owned = [k + 2.0 for k in range(k0, k1)]
        if path == 0 and rank == 0:
            owned[0] = Nz - 1 + 2.0
        assert r["owned"] == owned, r
        # ghosts hold the neighbours' new boundary planes (the outer ghosts of the end ranks are unused)
        if rank > 0:
            assert r["lo_ghost"] == (k0 - 1) + 2.0
        if rank < world - 1:
            assert r["hi_ghost"] == k1 + 2.0
