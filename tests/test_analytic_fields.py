"""The reference's own (unimplemented) test plan, /root/reference/tests/README.md:4-17: kernels on uniform, linear,
parabolic and sinusoidal fields; diffusion likewise.  Run on the CPU oracle -- the CUDA operators are bit-exact
against it (tests/test_gpu_parity.py), so what holds here holds for them.  These checks are independent of the
reference's code: they pin signs, stencil directions and coefficients to the equations."""
import numpy as np
import pytest

DT, D = 1e-3, 0.05
H = (0.125, 0.25, 0.5)  # dx, dy, dz: powers of two, so the test fields below are exact in fp32
GAMMA = 5.0 / 3.0


def uniform_state(Nx, Ny, Nz):
    Q = np.empty((8, Nz, Nx, Ny), np.float32)
    for v, val in enumerate((1.25, 0.5, -0.25, 0.75, 0.5, 0.25, -0.5, 4.0)):
        Q[v] = val
    return Q


def interior(a, lo=1):
    return a[:, lo:-1, lo:-1, lo:-1]


@pytest.mark.parametrize("path", [0, 1])
def test_uniform_state_is_a_fixed_point(O, oracle_mod, path):
    """All flux differences and the Laplacian vanish: predictor and corrector return the state, bit for bit."""
    Q = uniform_state(10, 9, 8)
    Qi = np.zeros_like(Q)
    O.predictor(Q, Qi, path, D, DT, *H)
    assert np.array_equal(interior(Qi), interior(Q))
    Q1 = Q.copy()
    O.corrector_volume(Q1, Qi, path, D, DT, *H)
    # path B: the wall faces predict with their outward flux set to zero (kernels_intvarbcs.cu:560-1110), and the
    # corrector's Laplacian of Qint reaches them from the first interior layer -> one more layer of margin
    m = 1 if path == 0 else 2
    inner = (slice(None), slice(m, -m), slice(m, -m), slice(m, -m))
    if path == 0:
        assert np.array_equal(Q1[inner], Q[inner])
    else:
        # quirk B-6 (live in the reference, reproduced): FluidAdvanceLocal takes rho(i-1) where the x-flux of rho at
        # i-1 (= rho vx) belongs (kernels_od.cu:120-345), so even a uniform state drifts: rho += (dt/2dx)(rho - rho vx)
        assert np.array_equal(Q1[inner][1:], Q[inner][1:])
        drift = 0.5 * (DT / H[0]) * (1.25 - 0.5)
        assert np.allclose(Q1[inner][0], 1.25 + drift, rtol=0, atol=2e-7)


@pytest.mark.parametrize("axis,mom", [(2, 1), (3, 2), (1, 3)])  # array axis (k, i, j order) <-> momentum component
def test_continuity_equation_uses_a_forward_then_a_backward_difference(O, oracle_mod, axis, mom):
    """rho_t + div(rho v) = 0 with rho v linear along one axis: the predictor takes (m(+1) - m)/h, the corrector
    (m'(0) - m'(-1))/h of the predicted momentum -- both equal the slope, so rho drops by dt * slope exactly."""
    Nx, Ny, Nz = 12, 10, 9
    Q = uniform_state(Nx, Ny, Nz)
    Q[4:7] = 0.0                                # no field: keeps the momentum equations out of the way of this check
    n = Q.shape[axis]
    slope = 0.5
    h = {2: H[0], 3: H[1], 1: H[2]}[axis]
    ramp = (slope * h * np.arange(n, dtype=np.float32)).reshape([-1 if a == axis else 1 for a in range(1, 4)])
    Q[mom] = Q[mom] + ramp.astype(np.float32)
    Qi = np.zeros_like(Q)
    O.predictor(Q, Qi, 0, 0.0, DT, *H)
    got = interior(Qi)[0] - interior(Q)[0]
    assert np.allclose(got, -DT * slope, rtol=0, atol=2e-7), (got.min(), got.max())
    # every OTHER direction contributes nothing to rho
    others = [m for m in (1, 2, 3) if m != mom]
    assert all(np.ptp(Q[m]) == 0 for m in others)


def test_diffusion_term_is_dt_D_laplacian(O, oracle_mod):
    """Path B predictor with D minus the same with D = 0 is dt * D * lap(Q) (diffusion.cu:8-19): zero on a linear
    field, dt * D * 2a/h^2 on a parabola a*n^2 along each axis, and -(2 - 2 cos(kh))/h^2 * q on a sine."""
    Nx, Ny, Nz = 16, 14, 12
    base = uniform_state(Nx, Ny, Nz)
    i = np.arange(Nx, dtype=np.float64).reshape(1, -1, 1)
    j = np.arange(Ny, dtype=np.float64).reshape(1, 1, -1)
    k = np.arange(Nz, dtype=np.float64).reshape(-1, 1, 1)

    def diffusion_part(field):
        Q = base.copy()
        Q[7] = (Q[7] + field).astype(np.float32)      # perturb the energy: it feeds only its own flux and the pressure
        a, b = np.zeros_like(Q), np.zeros_like(Q)
        O.predictor(Q, a, 1, D, DT, *H)
        O.predictor(Q, b, 1, 0.0, DT, *H)
        return interior(a)[7].astype(np.float64) - interior(b)[7].astype(np.float64), Q[7]

    lin, _ = diffusion_part(0.25 * i + 0.5 * j - 0.125 * k)
    assert np.abs(lin).max() <= 5e-7                     # rounding of the two runs only
    for axis_field, h in ((0.03125 * i * i, H[0]), (0.03125 * j * j, H[1]), (0.03125 * k * k, H[2])):
        par, _ = diffusion_part(axis_field)
        assert np.allclose(par, DT * D * 2 * 0.03125 / h ** 2, rtol=0, atol=6e-7)
    kx = 2 * np.pi / 8                                   # 8 points per wavelength along x
    sine, q = diffusion_part(0.5 * np.sin(kx * i) + 0 * j + 0 * k)
    pert = interior(q[None])[0].astype(np.float64) - 4.0
    want = -DT * D * (2 - 2 * np.cos(kx)) / H[0] ** 2 * pert
    assert np.allclose(sine, want, rtol=0, atol=1e-6)


def test_pressure_gradient_accelerates_down_the_gradient(O, oracle_mod):
    """A fluid at rest without field and with e linear in x: d(rho vx)/dt = -dp/dx with the reference's pressure
    p = (gamma - 1) e for a state at rest (helper_functions.cu:17-19)."""
    Nx, Ny, Nz = 12, 10, 8
    Q = np.zeros((8, Nz, Nx, Ny), np.float32)
    Q[0] = 1.0
    slope = 0.5
    Q[7] = (2.0 + slope * H[0] * np.arange(Nx, dtype=np.float32)).reshape(1, -1, 1)
    Qi = np.zeros_like(Q)
    O.predictor(Q, Qi, 0, 0.0, DT, *H)
    assert np.allclose(interior(Qi)[1], -DT * (GAMMA - 1) * slope, rtol=0, atol=2e-7)
    assert not interior(Qi)[2].any() and not interior(Qi)[3].any()
    assert np.array_equal(interior(Qi)[0], interior(Q)[0])


def test_entropy_wave_converges_at_second_order(O, oracle_mod):
    """"Kernels vs problem size" (tests/README.md:6-9): a density wave carried along the periodic z axis by a uniform
    flow (constant pressure, no field) is an exact solution rho(z - u t) of the reference's equations too (its
    kinetic-energy convention, B-1, is used consistently).  Lax-Wendroff must converge on it at second order: halving
    dz at fixed CFL number divides the error by ~4.  A flow ALONG the rigid walls is unstable in the reference's wall
    treatment (the wall cells blow up within ten steps) and its one-sided periodic seam (plane 0 is a copy of plane
    Nz-1, kernels_fluidbcs.cu:498-510) sheds plane-to-plane noise that grows with the step count, so the error is
    measured away from both: a cross-section wider than the walls' reach in these few steps, and only the planes far
    from the seam (measured ratio 33 -> 65 planes: 4.08)."""
    om = oracle_mod
    u0, p0, eps, Lz, T = 1.0, 1.0, 0.05, 1.0, 0.1
    errs = []
    for nz in (33, 65):
        nsteps = 5 * (nz - 1) // 16             # dt = T / nsteps: (u + c_s) dt/dz = 0.73 at both resolutions
        Nx = Ny = 2 * nsteps + 6
        dz = Lz / (nz - 1)                      # period of the reference's z axis: Nz - 1 intervals (plane Nz-1 == plane 0)
        dx = dy = 2.0
        z = dz * np.arange(nz)
        rho = 1.0 + eps * np.sin(2 * np.pi * z / Lz)
        Q = np.zeros((8, nz, Nx, Ny), np.float32)
        Q[0] = rho.reshape(-1, 1, 1)
        Q[3] = (rho * u0).reshape(-1, 1, 1)
        Q[7] = (p0 / (GAMMA - 1) + rho * u0 * u0).reshape(-1, 1, 1)   # e = p/(gamma-1) + KE with KE = rho v^2 (B-1)
        dt = T / nsteps
        Qi = np.zeros_like(Q)
        O.prime(Q, Qi, om.PATH_A, 0.0, dt, dx, dy, dz)
        O.steps(Q, Qi, om.PATH_A, nsteps, 0.0, dt, dx, dy, dz)
        exact = 1.0 + eps * np.sin(2 * np.pi * (z - u0 * T) / Lz)
        col = Q[0, :, Nx // 2, Ny // 2].astype(np.float64)
        w = slice(nsteps + 2, nz - nsteps - 2)
        errs.append(np.sqrt(np.mean((col[w] - exact[w]) ** 2)))
    assert errs[0] < 0.2 * eps                  # the coarse run already tracks the wave
    assert 3.0 < errs[0] / errs[1] < 5.5, errs  # second order
