"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(libimhd_b200.so); the CPU oracle is only the checker.

Bars (BASELINE.json north_star):
  * parity-granular operators: BIT-EXACT against the oracle (they keep the reference's rounding points)
  * fused hot path: per-variable normalised L-inf  max|new-ref| / max|ref|  <= 1e-5 after 100 steps (fp32)
  * boundary / indexing work (cell sets, copies, wall constants): bit-exact
"""
import json
import os
import threading

import numpy as np
import pytest

from conftest import BOUNDS, bits_equal, make_case
from test_oracle_golden import changed, random_state

pytestmark = pytest.mark.gpu

TOL = 1e-5  # normalised L-inf after 100 steps, fp32 (BASELINE.json north_star)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))
DT = 1e-4
D_B = 0.01


@pytest.fixture(scope="module")
def torch():
    import torch as t

    if not t.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return t


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def paths(om):
    return ((om.PATH_A, 0.0), (om.PATH_B, D_B))


# ------------------------------------------------------------------------------------------------------
# initial conditions and grids (lib/on-device/initialize_od.cu)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims", [(16, 12, 10), (64, 64, 64), (50, 34, 21)])
def test_grids_and_screwpinch_bit_exact(imhd, torch, O, oracle_mod, dims):
    Nx, Ny, Nz = dims
    g, _, Q0 = make_case(O, oracle_mod, *dims)
    gx, gy, gz = imhd.ops.init_grids(BOUNDS, *dims)
    for a, b in zip((gx, gy, gz), g):
        assert bits_equal(a.cpu().numpy(), b)
    Q = imhd.ops.init_screwpinch_stride(1.0, gx, gy, gz).cpu().numpy()
    assert bits_equal(Q, Q0)


def test_screwpinch_matches_reference_csv(imhd, torch):
    ref = np.loadtxt(os.path.join(GOLD, "ref_var0_rhovz_plane.csv"), delimiter=",", comments="#")
    gx, gy, gz = imhd.ops.init_grids(BOUNDS, 64, 64, 64)
    rhovz = imhd.ops.init_screwpinch_stride(1.0, gx, gy, gz)[3].cpu().numpy()
    assert int((rhovz[0] != 0).sum()) == 392
    np.testing.assert_allclose(rhovz[0], ref, rtol=0, atol=1e-6)
    assert all(np.array_equal(rhovz[0], rhovz[k]) for k in range(64))


def test_bennett_vortex_ic(imhd, torch, O, oracle_mod):
    dims = (40, 36, 20)
    g, _, Q0 = make_case(O, oracle_mod, *dims, ic="bennett")
    gx, gy, gz = imhd.ops.init_grids(BOUNDS, *dims)
    Q = imhd.ops.init_cubic_bennett_vortex_m0(2.0, 0.5, gx, gy, gz).cpu().numpy()
    # logf / cosf of CUDA and glibc may differ in the last bit: 2 ulp of O(1) values
    assert np.array_equal(Q == 0, Q0 == 0)
    np.testing.assert_allclose(Q, Q0, rtol=0, atol=5e-7)


@pytest.mark.parametrize("dims", [(40, 36, 20), (24, 18, 5)])
def test_other_initial_conditions(imhd, torch, O, oracle_mod, dims):
    """CubicBennettVortex, ZPinch, ScrewPinch (initialize_od.cu:59, 347, 207) against the oracle."""
    g = O.init_grids(BOUNDS, *dims)
    gx, gy, gz = imhd.ops.init_grids(BOUNDS, *dims)
    for coeff in (0.25, 0.4):
        assert bits_equal(imhd.ops.init_zpinch(coeff, gx, gy, gz).cpu().numpy(), O.zpinch(coeff, *g))
        pre = random_state(*dims, seed=5)
        Q = imhd.ops.init_screwpinch(1.0, coeff, gx, gy, gz, prefill=dev(torch, pre)).cpu().numpy()
        assert bits_equal(Q, O.screwpinch(1.0, coeff, *g, prefill=pre))
    # logf of CUDA and glibc may differ in the last bit
    Q, Q0 = imhd.ops.init_cubic_bennett_vortex(gx, gy, gz).cpu().numpy(), O.cubic_bennett_vortex(*g)
    assert np.array_equal(Q == 0, Q0 == 0)
    np.testing.assert_allclose(Q, Q0, rtol=0, atol=5e-7)


def test_registry_initializers_run_the_named_kernel(imhd, torch, O, oracle_mod):
    dims = (24, 18, 6)
    g = O.init_grids(BOUNDS, *dims)
    want = {("screwpinch-stride", (1.0,)): O.screwpinch_stride(1.0, *g),
            ("screwpinch", (1.0, 0.3)): O.screwpinch(1.0, 0.3, *g),  # the context clears the buffer first
            ("zpinch", (0.25,)): O.zpinch(0.25, *g)}
    with imhd.ops.Context(*dims) as ctx:
        with pytest.raises(imhd.ImhdError, match="before imhd_ctx_init_grids"):
            ctx.initialize("zpinch", 0.25)
        ctx.init_grids(*BOUNDS)
        for (key, params), Q0 in want.items():
            ctx.initialize(key, *params)
            assert bits_equal(ctx.get_state(), Q0), key
        for key, params in (("cubic-bennett-vortex", ()), ("cubic-bennett-vortex-m0", (2.0, 0.5))):
            ctx.initialize(key, *params)
            Q0 = O.cubic_bennett_vortex(*g) if not params else O.cubic_bennett_vortex_m0(2.0, 0.5, *g)
            np.testing.assert_allclose(ctx.get_state(), Q0, rtol=0, atol=5e-7)
        with pytest.raises(imhd.ImhdError, match="Unknown simulation type: orszag-tang"):
            ctx.initialize("orszag-tang")
        with pytest.raises(imhd.ImhdError, match="takes 2 parameter"):
            ctx.initialize("screwpinch", 1.0)
        # a registry-selected job: keys -> path -> prime -> step
        path = imhd.ops.registry_resolve_path("fluidadvancelocal", "corrector_advance-stride")
        ctx.initialize("screwpinch-stride", 1.0)
        ctx.prime(path, D_B, DT)
        ctx.step(3)
        Q = ctx.get_state()
    d = tuple(float(oracle_mod.grid_spacing(BOUNDS[2 * a], BOUNDS[2 * a + 1], n)) for a, n in enumerate(dims))
    Qo, Qi = O.screwpinch_stride(1.0, *g), np.zeros((8, dims[2], dims[0], dims[1]), np.float32)
    O.prime(Qo, Qi, path, D_B, DT, *d)
    O.steps(Qo, Qi, path, 3, D_B, DT, *d)
    assert oracle_mod.normalised_linf(Q, Qo).max() <= 1e-6


# ------------------------------------------------------------------------------------------------------
# parity-granular operators: bit-exact
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims", [(16, 12, 10), (10, 8, 7), (46, 38, 19)])
@pytest.mark.parametrize("ic", ["screwpinch", "bennett", "random"])
def test_granular_operators_bit_exact(imhd, torch, O, oracle_mod, dims, ic):
    ops, om = imhd.ops, oracle_mod
    g, (dx, dy, dz), Q0 = make_case(O, om, *dims, ic="bennett" if ic == "random" else ic)
    if ic == "random":
        Q0 = random_state(*dims)
    for path, D in paths(om):
        dt = 1e-3
        qo, io = Q0.copy(), np.full_like(Q0, np.nan)
        O.prime(qo, io, path, D, dt, dx, dy, dz)
        Q, Qi = dev(torch, Q0), torch.full(Q0.shape, float("nan"), device="cuda")
        if path == om.PATH_A:
            ops.initial_bcs(Q)
        ops.predictor(Q, Qi, path, D, dt, dx, dy, dz)
        assert bits_equal(Q.cpu().numpy(), qo) and bits_equal(Qi.cpu().numpy(), io)
        for _ in range(3):
            O.corrector_volume(qo, io, path, D, dt, dx, dy, dz)
            ops.corrector(Q, Qi, path, D, dt, dx, dy, dz)
            assert bits_equal(Q.cpu().numpy(), qo), "corrector"
            O.fluid_bcs(qo, io, path, D, dt, dx, dy, dz)
            ops.fluid_bcs(Q, Qi, path, D, dt, dx, dy, dz)
            assert bits_equal(Q.cpu().numpy(), qo), "fluid boundary pass"
            O.predictor(qo, io, path, D, dt, dx, dy, dz)
            ops.predictor(Q, Qi, path, D, dt, dx, dy, dz)
            assert bits_equal(Qi.cpu().numpy(), io), "predictor + Qint boundary passes"


@pytest.mark.parametrize("ic", ["screwpinch", "bennett"])
@pytest.mark.parametrize("tag", ["A", "B"])
def test_granular_reproduces_golden_fixtures(imhd, torch, ic, tag):
    name = f"small_{ic}_path{tag}.npz"
    gold, meta = np.load(os.path.join(GOLD, name)), MANIFEST["cases"][name]
    path = imhd.PATH_A if tag == "A" else imhd.PATH_B
    with imhd.Context(*meta["dims"]) as c:
        c.set_state(gold["Q_ic"])
        c.set_spacing(meta["dx"], meta["dy"], meta["dz"])
        c.prime(path, meta["D"], MANIFEST["dt"])
        assert bits_equal(c.get_state(), gold["Q_primed"])
        done = 0
        for n in (1, 2, 10):
            c.step_granular(n - done)
            done = n
            assert bits_equal(c.get_state(), gold[f"Q_step{n}"])


# ------------------------------------------------------------------------------------------------------
# fused hot path
# ------------------------------------------------------------------------------------------------------
def run_fused(imhd, Q0, path, D, dt, d, nsteps, chunk=0):
    Nz, Nx, Ny = Q0.shape[1:]
    imhd._lib.load().imhd_set_chunk(chunk)
    try:
        with imhd.Context(Nx, Ny, Nz) as c:
            c.set_state(Q0)
            c.set_spacing(*d)
            c.prime(path, D, dt)
            c.step(nsteps)
            return c.get_state()
    finally:
        imhd._lib.load().imhd_set_chunk(0)


@pytest.mark.parametrize("ic", ["screwpinch", "bennett"])
@pytest.mark.parametrize("tag", ["A", "B"])
def test_fused_c1_100_steps_within_tolerance(imhd, torch, O, oracle_mod, ic, tag):
    """BASELINE.json configs[0]: 64x64x128, 100 steps, per-variable normalised L-inf <= 1e-5."""
    om = oracle_mod
    path, D = (om.PATH_A, 0.0) if tag == "A" else (om.PATH_B, D_B)
    g, d, Q0 = make_case(O, om, 64, 64, 128, ic=ic)
    qo, io = Q0.copy(), np.zeros_like(Q0)
    O.prime(qo, io, path, D, DT, *d)
    O.steps(qo, io, path, 100, D, DT, *d)
    if ic == "screwpinch":  # the committed fixture, generated from the reference's own kernels
        meta = MANIFEST["cases"][f"c1_path{tag}.npz"]
        gold = np.load(os.path.join(GOLD, f"c1_path{tag}.npz"))
        assert bits_equal(qo[:, ::4, ::4, ::4].copy(), gold["sample"]) and meta["steps"] == 100
    Q = run_fused(imhd, Q0, path, D, DT, d, 100)
    assert np.isfinite(Q).all()
    err = om.normalised_linf(Q, qo)
    print(f"\n[{ic} path {tag}] normalised L-inf after 100 steps:", " ".join(f"{e:.2e}" for e in err))
    assert (err <= TOL).all(), err


@pytest.mark.parametrize("dims", [(16, 12, 10), (10, 8, 7), (66, 35, 23), (30, 62, 40)])
def test_fused_small_and_ragged_grids(imhd, torch, O, oracle_mod, dims):
    """Tiles larger than the domain, partial edge tiles, odd plane counts; random state -> every cell active."""
    om = oracle_mod
    g, d, _ = make_case(O, om, *dims)
    Q0 = random_state(*dims)
    for path, D in paths(om):
        qo, io = Q0.copy(), np.zeros_like(Q0)
        O.prime(qo, io, path, D, 1e-3, *d)
        O.steps(qo, io, path, 5, D, 1e-3, *d)
        Q = run_fused(imhd, Q0, path, D, 1e-3, d, 5)
        err = om.normalised_linf(Q, qo)
        assert (err <= 2e-6).all(), (path, err)


def test_fused_cell_sets_bit_exact(imhd, torch, O, oracle_mod):
    """Which cells change, the periodic copies and the wall constants must match the reference bit for bit."""
    om = oracle_mod
    dims = (34, 30, 17)
    Nx, Ny, Nz = dims
    g, d, _ = make_case(O, om, *dims)
    Q0 = random_state(*dims)
    for path, D in paths(om):
        qo, io = Q0.copy(), np.zeros_like(Q0)
        O.prime(qo, io, path, D, 1e-3, *d)
        P = qo.copy()
        O.steps(qo, io, path, 1, D, 1e-3, *d)
        Q = run_fused(imhd, Q0, path, D, 1e-3, d, 1)
        assert np.array_equal(changed(Q, P).any(0), changed(qo, P).any(0)), "set of updated cells"
        untouched = ~changed(qo, P).any(0)
        assert bits_equal(Q[:, untouched], P[:, untouched]), "cells no pass touches are carried over unchanged"
        if path == om.PATH_A:
            assert bits_equal(Q[:, 0], Q[:, -1]), "PBCs"
        else:
            for i in (0, -1):  # wall constants
                assert bits_equal(Q[:, 0, i, :], qo[:, 0, i, :])
            assert bits_equal(Q[:, -1, -1, -1], qo[:, -1, -1, -1])
            # the k=0 face formula is the exact one, but its Qint inputs come from the fast recipe
            assert (om.normalised_linf(Q[:, :1], qo[:, :1]) <= 1e-6).all()


@pytest.mark.parametrize("dims,variant", [((40, 36, 29), 0), ((36, 64, 29), 4)])  # 4: remainder-strip kernel forced on
def test_plane_range_launches_compose(imhd, torch, O, oracle_mod, dims, variant):
    """imhd_step_fused_planes over any split of the owned planes == one imhd_step_fused call, bit for bit."""
    om, ops = oracle_mod, imhd.ops
    Nx, Ny, Nz = dims
    g, d, Q0 = make_case(O, om, *dims, ic="bennett")
    lib = imhd._lib.load()
    for path, D in paths(om):
        lib.imhd_set_kernel_variant(variant)
        Qin = dev(torch, Q0)
        ref, out = torch.empty_like(Qin), torch.full_like(Qin, float("nan"))
        s = ops.make_slab(Nx, Ny, Nz, path, D, DT, *d)
        q0 = ops.qint_plane(Qin, 0, s)
        qw = ops.qint_plane(Qin, Nz - 2, s) if path == om.PATH_B else None
        ops.step_fused(Qin, ref, q0, q0, qw, s)
        for a, b in ((20, Nz), (0, 5), (5, 6), (6, 20)):   # out of order on purpose
            ops.step_fused_planes(Qin, out, q0, q0, qw, s, a, b)
        # both ends in ONE launch (two plane ranges), then the interior: what the slab loop does; degenerate splits too
        out2 = torch.full_like(Qin, float("nan"))
        ops.step_fused_ends(Qin, out2, q0, q0, qw, s, 0, 4, Nz - 4, Nz)
        ops.step_fused_ends(Qin, out2, q0, q0, qw, s, 4, 9, 9, Nz - 4)        # empty gap: one range
        out3 = torch.full_like(Qin, float("nan"))
        ops.step_fused_ends(Qin, out3, q0, q0, qw, s, 0, 1, Nz - 1, Nz)       # the two planes with their own rules only
        ops.step_fused_ends(Qin, out3, q0, q0, qw, s, 1, 7, 11, Nz - 1)
        ops.step_fused_planes(Qin, out3, q0, q0, qw, s, 7, 11)
        lib.imhd_set_kernel_variant(0)
        assert bits_equal(out.cpu().numpy(), ref.cpu().numpy()), path
        assert bits_equal(out2.cpu().numpy(), ref.cpu().numpy()), path
        assert bits_equal(out3.cpu().numpy(), ref.cpu().numpy()), path


def test_fused_is_chunking_independent(imhd, torch, O, oracle_mod):
    """z-chunks re-derive their predictor planes; any chunk length must give the same bits."""
    om = oracle_mod
    dims = (40, 36, 37)
    g, d, Q0 = make_case(O, om, *dims, ic="bennett")
    for path, D in paths(om):
        ref = run_fused(imhd, Q0, path, D, DT, d, 6, chunk=0)
        for chunk in (2, 3, 5, 36, 37):
            assert bits_equal(run_fused(imhd, Q0, path, D, DT, d, 6, chunk=chunk), ref), (path, chunk)


def test_tma_and_plain_load_variants_give_the_same_bits(imhd, torch, O, oracle_mod):
    """The hot kernel stages Q tiles with TMA (cp.async.bulk.tensor); grids with Ny % 4 != 0 use plain loads."""
    om = oracle_mod
    lib = imhd._lib.load()
    dims = (52, 44, 33)
    g, d, Q0 = make_case(O, om, *dims, ic="bennett")
    for path, D in paths(om):
        a = run_fused(imhd, Q0, path, D, DT, d, 6)
        lib.imhd_set_kernel_variant(1)
        try:
            b = run_fused(imhd, Q0, path, D, DT, d, 6)
        finally:
            lib.imhd_set_kernel_variant(0)
        assert bits_equal(a, b), path


@pytest.mark.parametrize("dims", [(40, 64, 21), (36, 96, 30), (64, 64, 40)])
def test_remainder_strip_kernel_gives_the_same_bits(imhd, torch, O, oracle_mod, dims):
    """When Ny leaves a few columns beyond the 30-wide (path A: 31-wide) tiles, a transposed strip kernel computes
    them instead of a whole extra tile column (304 = 10 x 30 + 2 + walls).  Same bits with and without it, under any
    chunking, and against the plain-load variant (which never uses the strip)."""
    om = oracle_mod
    lib = imhd._lib.load()
    g, d, Q0 = make_case(O, om, *dims, ic="bennett")
    Q0 = Q0 + 0.01 * random_state(*dims, seed=3)   # break the symmetry of the analytic IC
    for path, D in paths(om):
        try:
            lib.imhd_set_kernel_variant(4)          # strip on (by default only plane ranges >= 64 use it)
            a = run_fused(imhd, Q0, path, D, DT, d, 5)
            for chunk in (2, 7):
                assert bits_equal(run_fused(imhd, Q0, path, D, DT, d, 5, chunk=chunk), a), (path, chunk)
            lib.imhd_set_kernel_variant(2)          # strip off: the hot kernel's extra tile column does the work
            b = run_fused(imhd, Q0, path, D, DT, d, 5)
            lib.imhd_set_kernel_variant(1)          # plain loads
            c = run_fused(imhd, Q0, path, D, DT, d, 5)
        finally:
            lib.imhd_set_kernel_variant(0)
        assert bits_equal(a, b) and bits_equal(a, c), path


@pytest.mark.parametrize("variant", [8, 16, 32, 4 | 256, 4 | 32])
def test_kernel_variants_give_the_same_bits(imhd, torch, O, oracle_mod, variant):
    """The marching kernels the library carries -- default: split-phase exchange barrier + warp-autonomous strip, the strip and
    the two z faces of path B running UNDER the marching kernel on a side stream; 8: everything on one stream; 16: one row
    per thread; 32: two rows per thread behind a block-wide barrier per plane; 256: the block-per-tile (lanes along i) strip --
    are the same arithmetic on the same values: identical bits, both pipelines, ragged grid with a remainder strip (Nz >= 66:
    the side stream is only used for launches of 64 planes or more)."""
    om = oracle_mod
    lib = imhd._lib.load()
    dims = (44, 64, 70)
    g, d, Q0 = make_case(O, om, *dims, ic="bennett")
    Q0 = Q0 + 0.01 * random_state(*dims, seed=5)
    for path, D in paths(om):
        try:
            lib.imhd_set_kernel_variant(4)          # default kernels, strip on
            a = run_fused(imhd, Q0, path, D, DT, d, 6)
            lib.imhd_set_kernel_variant(variant)
            b = run_fused(imhd, Q0, path, D, DT, d, 6)
        finally:
            lib.imhd_set_kernel_variant(0)
        assert bits_equal(a, b), (path, variant)


def test_fused_matches_granular_on_device(imhd, torch, O, oracle_mod):
    """Two independent CUDA implementations of the same step agree to rounding."""
    om = oracle_mod
    g, d, Q0 = make_case(O, om, 48, 40, 30, ic="bennett")
    for path, D in paths(om):
        with imhd.Context(48, 40, 30) as a, imhd.Context(48, 40, 30) as b:
            for c in (a, b):
                c.set_state(Q0); c.set_spacing(*d); c.prime(path, D, DT)
            a.step(20); b.step_granular(20)
            assert (om.normalised_linf(a.get_state(), b.get_state()) <= 2e-6).all()


def test_run_host_is_the_same_job(imhd, torch, O, oracle_mod):
    om = oracle_mod
    g, d, Q0 = make_case(O, om, 32, 28, 16)
    out = np.empty_like(Q0)
    with imhd.Context(32, 28, 16) as c:
        c.run_host(Q0, out, om.PATH_B, D_B, DT, *d, 7)
    assert bits_equal(out, run_fused(imhd, Q0, om.PATH_B, D_B, DT, d, 7))


# ------------------------------------------------------------------------------------------------------
# z-slab decomposition on one GPU: P in-process "ranks" with a thread ring
# ------------------------------------------------------------------------------------------------------
class ThreadRing:
    def __init__(self, world):
        self.world, self.box, self.bar = world, {}, threading.Barrier(world)

    def comm(self, rank):
        ring = self

        class C:
            pass

        c = C()
        c.rank, c.world = rank, self.world

        def ring_exchange(send_up, send_down, recv_from_down, recv_from_up, up, down):
            ring.box[(rank, "up")], ring.box[(rank, "down")] = send_up, send_down
            torch_sync()   # the sender's kernels (its own, possibly non-blocking, stream) finish before a peer copies
            ring.bar.wait()
            recv_from_down.copy_(ring.box[(down, "up")])
            recv_from_up.copy_(ring.box[(up, "down")])
            torch_sync()
            ring.bar.wait()

        c.ring_exchange = ring_exchange
        return c


def torch_sync():
    import torch as t

    t.cuda.synchronize()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("Nz", [26, 62])  # 62: slabs long enough for the overlapped (ends-first, side-stream) schedule
def test_slab_decomposition_is_bit_identical(imhd, torch, O, oracle_mod, world, Nz):
    """results must be independent of the number of slabs bit for bit."""
    om = oracle_mod
    slab = __import__("importlib").import_module("imhd-cuda_b200.slab")
    dims = (36, 30, Nz)
    Nx, Ny, Nz = dims
    g, d, Q0 = make_case(O, om, *dims, ic="bennett")
    for path, D in paths(om):
        Qp = Q0.copy()
        if path == om.PATH_A:  # the one-shot initial boundary pass is applied to the global state
            O.wall_bcs_leftright(Qp); O.pbcs(Qp)
        ce = imhd.ops.wall_energy_fixed_point(float(Qp[7, 0, -1, -1]), Nx)
        ref = run_fused(imhd, Q0, path, D, DT, d, 5)
        ring, out, errs = ThreadRing(world), [None] * world, []

        def work(r):
            try:
                s = slab.SlabSolver(Nx, Ny, Nz, path, D, DT, *d, comm=ring.comm(r), corner_e=ce)
                assert s.overlap == (Nz == 62)
                s.load_global(Qp)
                s.step(5)
                torch.cuda.synchronize()
                out[r] = s.state.cpu().numpy()
            except Exception as e:  # noqa: BLE001
                errs.append(e)
                ring.bar.abort()

        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        assert bits_equal(np.concatenate(out, axis=1), ref), f"path {path}, {world} slabs"


# ------------------------------------------------------------------------------------------------------
# production size (BASELINE.json configs[1]): size-independent properties + a short direct comparison
# ------------------------------------------------------------------------------------------------------
def test_production_size_properties(imhd, torch, O, oracle_mod):
    """304x304x592 with diffusion on.  (1) 2 steps against the oracle; (2) the screw pinch is z-invariant
    and z-information travels one plane per step, so after n steps planes far from both z ends must still be
    bit-identical to each other."""
    om = oracle_mod
    Nx, Ny, Nz = 304, 304, 592
    g, d, Q0 = make_case(O, om, Nx, Ny, Nz)
    qo, io = Q0.copy(), np.zeros_like(Q0)
    O.prime(qo, io, om.PATH_B, D_B, DT, *d)
    O.steps(qo, io, om.PATH_B, 2, D_B, DT, *d)
    with imhd.Context(Nx, Ny, Nz) as c:
        c.set_state(Q0); c.set_spacing(*d); c.prime(om.PATH_B, D_B, DT)
        c.step(2)
        Q = c.get_state()
        err = om.normalised_linf(Q, qo)
        assert (err <= 1e-6).all(), err
        c.step(8)
        Q = c.get_state()
    assert np.isfinite(Q).all()
    n = 10
    mid = Q[:, 2 * n + 2]
    for k in (2 * n + 3, Nz // 2, Nz - 2 * n - 3):
        assert bits_equal(Q[:, k], mid), f"plane {k} lost z-invariance"


def test_bench_workload_spacing_100_steps(imhd, torch, O, oracle_mod):
    """The spacing bench.py runs at EVERY N (weak scaling extends z at constant dz: dx = dy = 2*3.14159/303,
    dz = 2*3.14159/591, dt = 1e-4, D = 0.01, diffusion number dt D (2/dx^2+2/dy^2+2/dz^2) = 0.027 < 1/2) on a thin
    box: 100 steps of the diffusion pipeline stay finite and inside 1e-5 of the oracle.  (Round 1 squeezed 592 N planes
    into the same box: at N = 8 that is 0.57 > 1/2 and the reference's own kernels blow up at step 14.)"""
    import importlib

    om = oracle_mod
    bench = importlib.import_module("bench")
    dx, dy, dz = bench.spacing(bench.NZ_PER_GPU)
    assert abs(bench.workload_config(8, "weak")["spacing"][2] - dz) < 1e-12      # constant dz at N = 8
    assert bench.workload_config(8, "weak")["diffusion_number"] < 0.5
    Nx, Ny, Nz = 72, 64, 20
    bounds = (-dx * (Nx - 1) / 2, dx * (Nx - 1) / 2, -dy * (Ny - 1) / 2, dy * (Ny - 1) / 2, 0.0, dz * (Nz - 1))
    g = O.init_grids(bounds, Nx, Ny, Nz)
    d = tuple(float(om.grid_spacing(bounds[2 * a], bounds[2 * a + 1], n)) for a, n in enumerate((Nx, Ny, Nz)))
    assert max(abs(a - b) / b for a, b in zip(d, (dx, dy, dz))) < 1e-5
    Q0 = O.screwpinch_stride(1.0, *g)
    qo, io = Q0.copy(), np.zeros_like(Q0)
    O.prime(qo, io, om.PATH_B, D_B, DT, *d)
    O.steps(qo, io, om.PATH_B, 100, D_B, DT, *d)
    Q = run_fused(imhd, Q0, om.PATH_B, D_B, DT, d, 100)
    assert np.isfinite(qo).all() and np.isfinite(Q).all()
    assert (om.normalised_linf(Q, qo) <= 1e-5).all()


def test_fused_path_converges_at_second_order(imhd, torch, O, oracle_mod):
    """Physics-level check of the product path through the C ABI: the advected density wave of
    tests/test_analytic_fields.py (an exact solution) converges at second order with the fused kernels too."""
    om = oracle_mod
    gamma = 5.0 / 3.0
    u0, p0, eps, Lz, T = 1.0, 1.0, 0.05, 1.0, 0.1
    errs = []
    for nz in (33, 65):
        nsteps = 5 * (nz - 1) // 16
        Nx = Ny = 2 * nsteps + 6
        dz, dx, dy = Lz / (nz - 1), 2.0, 2.0
        z = dz * np.arange(nz)
        rho = 1.0 + eps * np.sin(2 * np.pi * z / Lz)
        Q = np.zeros((8, nz, Nx, Ny), np.float32)
        Q[0] = rho.reshape(-1, 1, 1)
        Q[3] = (rho * u0).reshape(-1, 1, 1)
        Q[7] = (p0 / (gamma - 1) + rho * u0 * u0).reshape(-1, 1, 1)
        dt = T / nsteps
        with imhd.Context(Nx, Ny, nz) as c:
            c.set_state(Q)
            c.set_spacing(dx, dy, dz)
            c.prime(om.PATH_A, 0.0, dt)
            c.step(nsteps)
            col = c.get_state()[0, :, Nx // 2, Ny // 2].astype(np.float64)
        exact = 1.0 + eps * np.sin(2 * np.pi * (z - u0 * T) / Lz)
        w = slice(nsteps + 2, nz - nsteps - 2)
        errs.append(np.sqrt(np.mean((col[w] - exact[w]) ** 2)))
    assert errs[0] < 0.2 * eps and 3.0 < errs[0] / errs[1] < 5.5, errs
