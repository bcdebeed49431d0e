"""CFL / stability scan: the reference's host scanner
(src/on-device/utils/compute_stability.cpp) restated in numpy (oracle/stability.py), the closed-form wave speeds the
CUDA scan uses, and the CUDA scan itself through the C ABI.  Parity of the restatement is UNPINNED (Eigen is not in
the image, the reference ships no fixture): it is anchored on theory -- see oracle/stability.py."""
import importlib

import numpy as np
import pytest

from conftest import BOUNDS, make_case
from test_oracle_golden import random_state

from oracle import stability as st


def physical_states(n, seed=0):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.2, 2.0, n)
    vel, B = rng.normal(0, 1, (n, 3)), rng.normal(0, 1, (n, 3))
    p = rng.uniform(0.1, 2.0, n)
    e = p / (st.GAMMA - 1) + 0.5 * rho * (vel ** 2).sum(1) + 0.5 * (B ** 2).sum(1)
    return np.concatenate([rho[:, None], rho[:, None] * vel, B, e[:, None]], 1).astype(np.float32)


def test_reference_x_jacobian_has_the_mhd_wave_spectrum():
    """compute_stability.cpp:183-274 (matrix A) is the textbook 8-wave Jacobian: its spectral radius is
    |u| + c_f.  This pins the closed form the CUDA kernel evaluates to the reference's own (correct) matrix."""
    U = physical_states(1500)
    sr, ws = st.spectral_radii(U), st.wave_speeds(U)
    assert np.max(np.abs(sr[:, 0] - ws[:, 0]) / ws[:, 0]) < 5e-6
    # full spectrum of one state: {0, u, u +- ca, u +- cs, u +- cf}
    A = st.jacobians(U[:1])[0][0].astype(np.float64)
    rho, u, Bx = U[0, 0], U[0, 1] / U[0, 0], U[0, 4]
    ev = np.sort(np.linalg.eigvals(A).real)
    ca = abs(Bx) / np.sqrt(rho)
    for lam in (0.0, u, u + ca, u - ca):
        assert np.min(np.abs(ev - lam)) < 1e-5
    assert ev[-1] - u == pytest.approx(u - ev[0], rel=1e-5)  # u +- c_f are symmetric about u


def test_reference_y_and_z_jacobians_carry_slips():
    """Quirk B-26: the y / z matrices of the reference are not the flux Jacobians (whole induction rows with the
    opposite sign, misplaced entries), so their spectra are not the wave speeds.  Recorded, not reproduced."""
    U = physical_states(400, seed=3)
    sr, ws = st.spectral_radii(U), st.wave_speeds(U)
    for d in (1, 2):
        assert np.median(np.abs(sr[:, d] - ws[:, d]) / ws[:, d]) > 0.02


def test_imaginary_speeds_use_the_complex_modulus():
    """p < 0: the slow pair becomes u +- i s; Eigen's abs() (:173) gives sqrt(u^2 + s^2)."""
    U = physical_states(300, seed=5)
    U[:, 7] *= 0.3  # drive the pressure negative
    assert ((st.GAMMA - 1) * (U[:, 7] - 0.5 * (U[:, 1:4] ** 2).sum(1) / U[:, 0] - 0.5 * (U[:, 4:7] ** 2).sum(1)) < 0).any()
    sr, ws = st.spectral_radii(U), st.wave_speeds(U)
    assert np.max(np.abs(sr[:, 0] - ws[:, 0]) / ws[:, 0]) < 2e-5


def test_scan_summary_and_slab_combination(O, oracle_mod):
    slab = importlib.import_module("imhd-cuda_b200.slab")
    dims = (16, 12, 10)
    _, d, Q = make_case(O, oracle_mod, *dims)
    lhs = st.wave_speed_lhs(Q, 1e-4, *d)
    s = st.scan(lhs, 1e-4)
    assert s["violations"] == 0 and 0 < s["max_lhs"] < 1e-2 and s["dt_new"] == pytest.approx(0.1 * 1e-4 / s["max_lhs"])
    big = st.scan(st.wave_speed_lhs(Q, 1.0, *d), 1.0)
    assert big["violations"] > 0 and big["max_lhs"] == pytest.approx(1e4 * s["max_lhs"], rel=1e-12)
    # splitting the domain into slabs and combining gives the global answer
    rows = []
    for k0, k1 in ((0, 3), (3, 7), (7, 10)):
        r = st.scan(lhs[k0:k1], 1e-4)
        i, j, k = r["argmax_ijk"]
        rows.append([r["max_lhs"], i, j, k + k0, r["violations"]])
    c = slab.combine_stability(rows, 1e-4)
    assert c["max_lhs"] == s["max_lhs"] and c["argmax_ijk"] == s["argmax_ijk"] and c["violations"] == 0
    assert c["dt_new"] == pytest.approx(s["dt_new"])
    assert slab.combine_stability([[0.5, 1, 1, 1, 2], [0.5, 0, 0, 9, 3]], 1.0)["argmax_ijk"] == (1, 1, 1)  # tie: first wins


# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def torch():
    import torch as t

    if not t.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return t


def check_scan(got, lhs, dt):
    want = st.scan(lhs, dt)
    assert got["max_lhs"] == pytest.approx(want["max_lhs"], rel=2e-6)
    k, i, j = np.unravel_index(np.argmax(np.nan_to_num(lhs, nan=-1.0)), lhs.shape)
    gi, gj, gk = got["argmax_ijk"]
    assert lhs[gk, gi, gj] >= want["max_lhs"] * (1 - 2e-6)          # the winner is a maximum up to fp32 rounding
    near = int((np.abs(np.nan_to_num(lhs, nan=0.0) - 1.0) < 5e-6).sum())  # cells too close to the threshold to call in fp32
    assert abs(got["violations"] - want["violations"]) <= near
    assert got["dt_new"] == pytest.approx(want["dt_new"], rel=2e-6)
    return want


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(16, 12, 10), (50, 34, 21), (9, 7, 5)])  # (9,7,5): plane % 4 != 0 -> scalar-load variant
def test_cuda_scan_matches_the_closed_form(imhd, torch, O, oracle_mod, dims):
    ops = imhd.ops
    Nx, Ny, Nz = dims
    _, d, Q0 = make_case(O, oracle_mod, *dims, ic="bennett")
    U = physical_states(Nx * Ny * Nz, seed=9).reshape(Nz, Nx, Ny, 8)
    Qr = np.ascontiguousarray(np.moveaxis(U, -1, 0))
    for Q, dt in ((Q0, 1e-4), (Q0, 0.05), (Qr, 0.01), (Qr, 0.05)):
        slab = ops.make_slab(Nx, Ny, Nz, 0, 0.0, dt, *d)
        got = ops.stability_scan(torch.from_numpy(Q).cuda(), slab)
        check_scan(got, st.wave_speed_lhs(Q, dt, *d), dt)
    assert got["violations"] > 0 or dims == (9, 7, 5)  # the coarse grid has no violation at dt = 0.05
    # context entry point
    with ops.Context(*dims) as ctx:
        ctx.set_state(Qr)
        ctx.set_spacing(*d)
        assert ctx.stability(0.05) == got


@pytest.mark.gpu
def test_cuda_scan_reference_quirks_mode(imhd, torch, O, oracle_mod):
    """IMHD_STABILITY_REFERENCE_QUIRKS: the spectral radii of the reference's OWN y and z matrices (slips B-26 included),
    Hessenberg + shifted QR per cell on the device, against the numpy restatement's LAPACK eigenvalues of the same
    matrices; the default mode keeps reporting the exact wave speeds."""
    ops = imhd.ops
    lib = imhd._lib.load()
    dims = (14, 12, 9)
    Nx, Ny, Nz = dims
    _, d, Q0 = make_case(O, oracle_mod, *dims, ic="bennett")
    U = physical_states(Nx * Ny * Nz, seed=4).reshape(Nz, Nx, Ny, 8)
    Qr = np.ascontiguousarray(np.moveaxis(U, -1, 0))
    try:
        lib.imhd_stability_mode(1)
        for Q, dt in ((Qr, 0.01), (Qr, 0.05), (Q0, 0.02)):
            slab = ops.make_slab(Nx, Ny, Nz, 0, 0.0, dt, *d)
            got = ops.stability_scan(torch.from_numpy(Q).cuda(), slab)
            lhs = st.reference_lhs(Q, dt, *d)
            want = st.scan(lhs, dt)
            assert got["max_lhs"] == pytest.approx(want["max_lhs"], rel=2e-4)   # fp32 LAPACK vs fp64 QR on fp32 matrices
            k, i, j = want["argmax_ijk"][2], want["argmax_ijk"][0], want["argmax_ijk"][1]
            gi, gj, gk = got["argmax_ijk"]
            assert lhs[gk, gi, gj] == pytest.approx(lhs[k, i, j], rel=2e-4)
            near = int((np.abs(np.nan_to_num(lhs, nan=0.0) - 1.0) < 5e-4).sum())
            assert abs(got["violations"] - want["violations"]) <= near
        # the slipped matrices do NOT have the wave spectrum: the two modes differ on generic states
        quirk = got["max_lhs"]
        lib.imhd_stability_mode(0)
        exact = ops.stability_scan(torch.from_numpy(Q0).cuda(), ops.make_slab(Nx, Ny, Nz, 0, 0.0, 0.02, *d))["max_lhs"]
        assert exact == pytest.approx(st.scan(st.wave_speed_lhs(Q0, 0.02, *d), 0.02)["max_lhs"], rel=2e-6)
        assert quirk != exact
    finally:
        lib.imhd_stability_mode(0)


@pytest.mark.gpu
def test_cuda_scan_edge_cases(imhd, torch, O, oracle_mod):
    ops = imhd.ops
    dims = (12, 8, 6)
    Nx, Ny, Nz = dims
    d = (0.1, 0.2, 0.3)
    U = physical_states(Nx * Ny * Nz, seed=2).reshape(Nz, Nx, Ny, 8)
    Q = np.ascontiguousarray(np.moveaxis(U, -1, 0))
    Q[:, 2, 3, 4] = 0.0           # rho = 0 -> NaN LHS: ignored (a NaN never passes `>= 1.0` in the reference either)
    Q[7, 4, 5, 1] *= 0.2          # p < 0 -> imaginary slow speed
    Q[:, 5, 11, 7] = Q[:, 0, 0, 0]  # duplicate of the first cell: ties resolve to the first in scan order
    dt = 0.02
    lhs = st.wave_speed_lhs(Q, dt, *d)
    assert np.isnan(lhs[2, 3, 4])
    got = ops.stability_scan(torch.from_numpy(Q).cuda(), ops.make_slab(Nx, Ny, Nz, 0, 0.0, dt, *d))
    check_scan(got, lhs, dt)
    # make the duplicated pair the maximum: the winner must be (0,0,0), not its copy at (11,7,5)
    Q2 = Q.copy()
    Q2[1:4, 0, 0, 0] *= 50.0
    Q2[7, 0, 0, 0] *= 2500.0
    Q2[:, 5, 11, 7] = Q2[:, 0, 0, 0]
    got = ops.stability_scan(torch.from_numpy(Q2).cuda(), ops.make_slab(Nx, Ny, Nz, 0, 0.0, dt, *d))
    assert got["argmax_ijk"] == (0, 0, 0)
    # a vacuum state at rest: LHS == 0 everywhere
    Z = np.zeros_like(Q)
    Z[0] = 0.01
    got = ops.stability_scan(torch.from_numpy(Z).cuda(), ops.make_slab(Nx, Ny, Nz, 0, 0.0, dt, *d))
    assert got == {"max_lhs": 0.0, "argmax_ijk": (0, 0, 0), "violations": 0, "dt_new": 0.0}


@pytest.mark.gpu
def test_cuda_scan_of_slabs_combines_to_the_global_scan(imhd, torch, O, oracle_mod):
    ops = imhd.ops
    slab_mod = importlib.import_module("imhd-cuda_b200.slab")
    dims = (20, 16, 23)
    Nx, Ny, Nz = dims
    d = (0.1, 0.1, 0.2)
    U = physical_states(Nx * Ny * Nz, seed=4).reshape(Nz, Nx, Ny, 8)
    Q = np.ascontiguousarray(np.moveaxis(U, -1, 0))
    dt = 0.03
    whole = ops.stability_scan(torch.from_numpy(Q).cuda(), ops.make_slab(Nx, Ny, Nz, 0, 0.0, dt, *d))
    for world in (2, 3, 5):
        rows = []
        for r in range(world):
            L = slab_mod.SlabLayout(Nz, world, r)
            buf = np.full((8, L.nzl + 2, Nx, Ny), np.nan, np.float32)   # ghost planes must not be scanned
            buf[:, 1:-1] = Q[:, L.k0:L.k1]
            s = ops.make_slab(Nx, Ny, Nz, 0, 0.0, dt, *d, k0=L.k0, nzl=L.nzl, ghosts=1)
            g = ops.stability_scan(torch.from_numpy(buf).cuda(), s)
            rows.append([g["max_lhs"], *g["argmax_ijk"], g["violations"]])
        c = slab_mod.combine_stability(rows, dt)
        assert c["max_lhs"] == whole["max_lhs"] and c["argmax_ijk"] == whole["argmax_ijk"]
        assert c["violations"] == whole["violations"]
        assert c["dt_new"] == pytest.approx(whole["dt_new"], rel=1e-6)


@pytest.mark.gpu
def test_adaptive_dt_loop_follows_the_scan_one_group_later(imhd):
    """imhd_ctx_step_adaptive: the scan taken before step g*every decides the dt of steps [(g+1)*every, (g+2)*every) as
    min(dt_max, cfl * dt_scan / max_lhs), the first group runs with the primed dt -- replayed here with synchronous scans and
    imhd_ctx_set_dt on a second context: the same dt sequence (exactly) and the same bits."""
    import torch

    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    dims, every, nsteps, cfl, dt0, dt_max = (48, 40, 32), 3, 11, np.float32(0.4), np.float32(1e-4), np.float32(2e-3)
    B = (-3.14159, 3.14159) * 3
    for path, D in ((imhd.PATH_B, 0.01), (imhd.PATH_A, 0.0)):
        with imhd.Context(*dims) as c:
            c.init_grids(*B); c.init_screwpinch_stride(1.0); c.prime(path, D, float(dt0))
            used = c.step_adaptive(nsteps, every=every, cfl_target=float(cfl), dt_max=float(dt_max))
            Qa = c.get_state()
        with imhd.Context(*dims) as r:
            r.init_grids(*B); r.init_screwpinch_stride(1.0); r.prime(path, D, float(dt0))
            dt, want, scans = dt0, [], []
            for it in range(nsteps):
                if it % every == 0:
                    if it >= every:   # the scan of the group before decides this group
                        dts, mx = scans[it // every - 1]
                        if mx > 0:
                            dt = min(np.float32(cfl * dts / np.float32(mx)), dt_max)
                            r.set_dt(float(dt))
                    scans.append((dt, r.stability(float(dt))["max_lhs"]))
                want.append(dt)
                r.step(1)
            Qb = r.get_state()
        assert np.array_equal(np.asarray(want, np.float32).view(np.uint32), used.view(np.uint32)), (path, want, used.tolist())
        assert len(set(used.tolist())) > 1 and used.max() <= dt_max      # the loop did adapt
        assert np.isfinite(Qa).all() and np.array_equal(Qa.view(np.uint32), Qb.view(np.uint32)), path
