"""CPU tests of the oracle (no GPU): the C restatement must reproduce, BIT FOR BIT, the fixtures that
tests/golden/make_golden.py generated from the reference's own kernels (oracle/_ref), the reference's
only shipped golden vector (debug/data/rhovz/var_0.csv), and -- where oracle/_ref is present -- the
reference itself on fresh cases."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import BOUNDS, bits_equal, make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))
DT = MANIFEST["dt"]


def _small(name):
    return np.load(os.path.join(GOLD, name)), MANIFEST["cases"][name]


@pytest.mark.parametrize("ic", ["screwpinch", "bennett"])
@pytest.mark.parametrize("tag", ["A", "B"])
def test_restatement_reproduces_small_goldens(O, oracle_mod, ic, tag):
    gold, meta = _small(f"small_{ic}_path{tag}.npz")
    Nx, Ny, Nz = meta["dims"]
    path = oracle_mod.PATH_A if tag == "A" else oracle_mod.PATH_B
    g, (dx, dy, dz), Q = make_case(O, oracle_mod, Nx, Ny, Nz, ic)
    assert (dx, dy, dz) == (meta["dx"], meta["dy"], meta["dz"])
    assert bits_equal(Q, gold["Q_ic"]), "initial condition differs from the reference's"
    Qi = np.zeros_like(Q)
    O.prime(Q, Qi, path, meta["D"], DT, dx, dy, dz)
    assert bits_equal(Q, gold["Q_primed"]) and bits_equal(Qi, gold["Qint_primed"])
    done = 0
    for n in (1, 2, 10):
        O.steps(Q, Qi, path, n - done, meta["D"], DT, dx, dy, dz)
        done = n
        assert bits_equal(Q, gold[f"Q_step{n}"]), f"Q after {n} steps"
        assert bits_equal(Qi, gold[f"Qint_step{n}"]), f"Qint after {n} steps"


@pytest.mark.parametrize("tag", ["A", "B"])
def test_restatement_reproduces_c1_100_steps(O, oracle_mod, tag):
    """BASELINE.json configs[0]: 64x64x128 screw pinch, 100 steps -- sha256 of the full state."""
    name = f"c1_path{tag}.npz"
    gold, meta = _small(name)
    Nx, Ny, Nz = meta["dims"]
    path = oracle_mod.PATH_A if tag == "A" else oracle_mod.PATH_B
    g, (dx, dy, dz), Q = make_case(O, oracle_mod, Nx, Ny, Nz)
    assert hashlib.sha256(Q.tobytes()).hexdigest() == meta["sha256_Q_ic"]
    Qi = np.zeros_like(Q)
    O.prime(Q, Qi, path, meta["D"], DT, dx, dy, dz)
    O.steps(Q, Qi, path, meta["steps"], meta["D"], DT, dx, dy, dz)
    assert np.isfinite(Q).all()
    assert bits_equal(Q[:, ::4, ::4, ::4].copy(), gold["sample"])
    assert hashlib.sha256(Q.tobytes()).hexdigest() == meta["sha256_Q"]
    np.testing.assert_array_equal(Q.reshape(8, -1).min(1), gold["vmin"])
    np.testing.assert_array_equal(Q.reshape(8, -1).max(1), gold["vmax"])


def test_initial_condition_matches_reference_csv(O, oracle_mod):
    """The reference's only golden vector: rho*v_z of ScrewPinchStride at t=0, 64^3 (printed to 6 digits)."""
    ref = np.loadtxt(os.path.join(GOLD, "ref_var0_rhovz_plane.csv"), delimiter=",", comments="#")
    g = O.init_grids(BOUNDS, 64, 64, 64)
    Q = O.screwpinch_stride(1.0, *g)
    rhovz = Q[3]
    assert all(np.array_equal(rhovz[0], rhovz[k]) for k in range(64)), "screw pinch IC must be z-invariant"
    assert int((rhovz[0] != 0).sum()) == MANIFEST["ref_var0"]["nonzero_per_plane"] == 392
    assert np.array_equal(rhovz[0] != 0, ref != 0), "support of the pinch differs"
    np.testing.assert_allclose(rhovz[0], ref, rtol=0, atol=1e-6)  # file holds 6 significant digits


@pytest.mark.parametrize("dims", [(10, 8, 7), (14, 18, 11)])  # even Nx, Ny: no grid point on the axis (0/0 in the IC)
@pytest.mark.parametrize("ic", ["screwpinch", "bennett"])
def test_restatement_bit_exact_vs_reference_kernels(O, R, oracle_mod, dims, ic):
    """Fresh cases (not in the fixtures) against the reference's kernels run on the host."""
    Nx, Ny, Nz = dims
    g, (dx, dy, dz), Q0 = make_case(R, oracle_mod, Nx, Ny, Nz, ic)
    go, _, Q0o = make_case(O, oracle_mod, Nx, Ny, Nz, ic)
    assert bits_equal(Q0, Q0o)
    for path, D in ((oracle_mod.PATH_A, 0.0), (oracle_mod.PATH_B, 0.05)):
        qo, qr = Q0.copy(), Q0.copy()
        io, ir = np.full_like(qo, np.nan), np.full_like(qo, np.nan)
        O.prime(qo, io, path, D, 1e-3, dx, dy, dz)
        R.prime(qr, ir, path, D, 1e-3, dx, dy, dz)
        assert bits_equal(qo, qr) and bits_equal(io, ir)  # also proves every Qint cell is written (no NaN left)
        for _ in range(4):
            O.steps(qo, io, path, 1, D, 1e-3, dx, dy, dz)
            R.steps(qr, ir, path, 1, D, 1e-3, dx, dy, dz)
            assert bits_equal(qo, qr) and bits_equal(io, ir)


def test_reference_boundary_megakernel_thread_subset_is_equivalent(R, oracle_mod):
    """ref_harness runs ComputeIntermediateVariablesBoundary on the non-duplicate threads only; the full
    one-thread-per-cell launch must give the same Qint."""
    Nx, Ny, Nz = 8, 9, 7
    g, (dx, dy, dz), Q = make_case(R, oracle_mod, Nx, Ny, Nz, "bennett")
    a, b = np.zeros_like(Q), np.zeros_like(Q)
    R.predictor(Q, a, oracle_mod.PATH_B, 0.05, 1e-3, dx, dy, dz, full=False)
    R.predictor(Q, b, oracle_mod.PATH_B, 0.05, 1e-3, dx, dy, dz, full=True)
    assert bits_equal(a, b)


# ---- cell sets / quirks the boundary passes must honour (bit-exact bookkeeping) --------------------------
def random_state(Nx, Ny, Nz, seed=7):
    """A smooth-free random state (every cell 'active'), so that 'cell X is (not) updated' shows in the bits."""
    rng = np.random.default_rng(seed)
    Q = rng.uniform(-0.1, 0.1, (8, Nz, Nx, Ny)).astype(np.float32)
    Q[0] = rng.uniform(0.5, 1.5, (Nz, Nx, Ny))
    Q[7] = rng.uniform(1.0, 2.0, (Nz, Nx, Ny))
    return Q


def changed(a, b):
    return a.view(np.uint32) != b.view(np.uint32)


def test_path_a_cell_sets(O, oracle_mod):
    Nx, Ny, Nz = 12, 10, 9
    g, (dx, dy, dz), _ = make_case(O, oracle_mod, Nx, Ny, Nz)
    Q0 = random_state(Nx, Ny, Nz)
    Q, Qi = Q0.copy(), np.zeros_like(Q0)
    O.prime(Q, Qi, oracle_mod.PATH_A, 0.0, 1e-3, dx, dy, dz)
    P = Q.copy()
    # init walls: j = 0 and Ny-1, 0<k<Nz-1 (kernels_fluidbcs.cu:436-464); i-walls are NOT imposed (B-11)
    assert (P[0, 1:-1, :, 0] == 1).all() and (P[0, 1:-1, :, -1] == 1).all()
    assert (P[1:7, 1:-1, :, 0] == 0).all() and (P[1:7, 1:-1, :, -1] == 0).all()
    assert bits_equal(P[:, 1:-1, :, 1:-1], Q0[:, 1:-1, :, 1:-1]), "nothing but the j-walls and plane 0 is touched at init"
    assert bits_equal(P[:, 0], P[:, -1]) and bits_equal(P[:, -1], Q0[:, -1])  # PBCs: plane 0 <- plane Nz-1
    O.steps(Q, Qi, oracle_mod.PATH_A, 1, 0.0, 1e-3, dx, dy, dz)
    ch = changed(Q, P).any(axis=0)  # (Nz, Nx, Ny): cells whose bits changed in one step
    assert not ch[1:, 0, :].any(), "i = 0 face is never updated (B-14)"
    assert not ch[1:, :, 0].any(), "j = 0 face is never updated (B-14)"
    assert ch[1:, 1:, 1:].all(), "every cell with i,j,k >= 1 is updated, far faces included (B-14)"
    assert bits_equal(Q[:, 0], Q[:, -1]), "PBCs copy back -> front every step"
    assert bits_equal(Qi[:, -1], Qi[:, 0]), "QintBdryPBCs copy front -> back"


def test_path_b_cell_sets(O, oracle_mod):
    Nx, Ny, Nz = 12, 10, 9
    g, (dx, dy, dz), _ = make_case(O, oracle_mod, Nx, Ny, Nz)
    Q0 = random_state(Nx, Ny, Nz)
    Q, Qi = Q0.copy(), np.zeros_like(Q0)
    O.prime(Q, Qi, oracle_mod.PATH_B, 0.05, 1e-3, dx, dy, dz)
    assert bits_equal(Q, Q0), "main.cu applies no boundary pass to the initial state"
    O.steps(Q, Qi, oracle_mod.PATH_B, 1, 0.05, 1e-3, dx, dy, dz)
    ch = changed(Q, Q0).any(axis=0)
    expect = np.zeros((Nz, Nx, Ny), bool)
    expect[1:-1, 1:-1, 1:-1] = True   # FluidAdvanceLocal: [1,N-2]^3
    expect[0, 1:-1, 1:-1] = True      # BoundaryConditions: k = 0 face interior
    expect[0, 0, :] = expect[0, -1, :] = True  # walls (0,j,0), (Nx-1,j,0) for all j; the j-walls are dead code (B-8)
    expect[-1, -1, -1] = True         # the 'PBC' copies only the column (Nx-1,Ny-1) (B-8)
    assert np.array_equal(ch, expect)
    for i in (0, -1):
        assert (Q[0, 0, i, :] == 1).all() and (Q[1:7, 0, i, :] == 0).all()
    assert bits_equal(Q[:, -1, -1, -1], Q[:, 0, -1, -1])


# ------------------------------------------------------------------------------------------------------
# the three initial conditions the shipped drivers keep commented out (initialize_od.cu:59, 207, 347)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims", [(10, 8, 7), (24, 18, 5)])
def test_other_initial_conditions_bit_exact_vs_reference_kernels(O, R, oracle_mod, dims):
    from conftest import BOUNDS

    g = O.init_grids(BOUNDS, *dims)
    assert bits_equal(O.cubic_bennett_vortex(*g), R.cubic_bennett_vortex(*g))
    for coeff in (0.25, 0.4):
        assert bits_equal(O.zpinch(coeff, *g), R.zpinch(coeff, *g))
        pre = random_state(*dims, seed=5)
        a, b = O.screwpinch(1.0, coeff, *g, prefill=pre), R.screwpinch(1.0, coeff, *g, prefill=pre)
        assert bits_equal(a, b)
        # outside the pinch ScrewPinch writes rho = 0.1 only (initialize_od.cu:237): the other seven keep the prefill
        outside = a[0] == np.float32(0.1)
        assert outside.any() and (~outside).any()
        assert bits_equal(a[1:][:, outside], pre[1:][:, outside])
        assert np.all(a[0][~outside] == 1.0)


def test_zpinch_and_bennett_are_z_invariant_equilibria(O, oracle_mod):
    from conftest import BOUNDS

    g = O.init_grids(BOUNDS, 20, 20, 6)
    for Q in (O.zpinch(0.25, *g), O.cubic_bennett_vortex(*g)):
        assert np.isfinite(Q).all()
        assert all(bits_equal(Q[:, k], Q[:, 0]) for k in range(Q.shape[1]))
        assert set(np.unique(Q[0])) == {np.float32(0.01), np.float32(1.0)}
        assert not Q[1].any() and not Q[2].any() and not Q[6].any()  # no in-plane flow, no axial field
