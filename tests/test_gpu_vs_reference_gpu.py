"""GPU parity against the reference ITSELF run on the B200: its unmodified kernels compiled by nvcc for sm_100
(oracle/_ref/libimhd_ref_gpu*.so, built by `make -C oracle refgpu` where /root/reference exists; the .so travels
to the GPU box).  Complements test_gpu_parity.py, whose checker is the reference compiled for the host.

  * -fmad=false build: rounding points == source.  The host build of the same sources must agree (pins the host
    oracle to the device semantics: CUDA libdevice pow / f64 division vs glibc), and so must our granular kernels.
  * stock build (default -fmad): the fused hot path must stay inside the 1e-5 bar against it after 100 steps at C1.
"""
import os

import numpy as np
import pytest

from conftest import bits_equal, make_case
from test_oracle_golden import random_state

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
DT, D_B, TOL = 1e-4, 0.01, 1e-5


@pytest.fixture(scope="module")
def torch():
    import torch as t

    if not t.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return t


def refgpu(oracle_mod, nofma):
    try:
        return oracle_mod.ReferenceGPU(nofma=nofma)
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref/libimhd_ref_gpu*.so not built (needs /root/reference at build time)")


def run_ref_gpu(torch, G, Q0, path, D, d, nsteps, geom):
    _, Nz, Nx, Ny = Q0.shape
    Q = torch.from_numpy(Q0.copy()).cuda()
    Qint = torch.zeros_like(Q)
    assert G.covers(geom, Nx, Ny, Nz)
    G.prime(Q.data_ptr(), Qint.data_ptr(), (Nx, Ny, Nz), path, D, DT, *d, geom)
    if nsteps:
        G.steps(Q.data_ptr(), Qint.data_ptr(), (Nx, Ny, Nz), path, nsteps, D, DT, *d, geom)
    torch.cuda.synchronize()
    return Q.cpu().numpy(), Qint.cpu().numpy()


@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("ic", ["screwpinch", "random"])
def test_reference_on_gpu_agrees_with_its_host_build_and_with_granular(imhd, torch, O, R, oracle_mod, tag, ic):
    G = refgpu(oracle_mod, nofma=True)
    path, D = (oracle_mod.PATH_A, 0.0) if tag == "A" else (oracle_mod.PATH_B, D_B)
    dims = (46, 38, 19)
    _, d, Q0 = make_case(O, oracle_mod, *dims)
    if ic == "random":
        Q0 = random_state(*dims, seed=11)
    geom = G.COVER_A if tag == "A" else G.COVER_B
    nsteps = 3
    Qg, Qig = run_ref_gpu(torch, G, Q0, path, D, d, nsteps, geom)

    # the host build of the same sources
    Qh, Qih = Q0.copy(), np.zeros_like(Q0)
    R.prime(Qh, Qih, path, D, DT, *d)
    R.steps(Qh, Qih, path, nsteps, D, DT, *d)
    eq = np.isclose(Qg, Qh, rtol=0, atol=0, equal_nan=True)
    err = oracle_mod.normalised_linf(np.nan_to_num(Qg), np.nan_to_num(Qh)).max()
    print(f"\nreference sm_100 (-fmad=false) vs reference host, path {tag} {ic}: bit-equal cells {eq.mean():.6f}, nLinf {err:.2e}")
    assert eq.all()   # exact, NaN == NaN (the random state drives a few cells non-finite in both builds alike)

    # our parity-granular operators through the C ABI
    with imhd.ops.Context(*dims) as ctx:
        ctx.set_state(Q0)
        ctx.set_spacing(*d)
        ctx.prime(path, D, DT)
        ctx.step_granular(nsteps)
        Qo = ctx.get_state()
    err2 = oracle_mod.normalised_linf(np.nan_to_num(Qo), np.nan_to_num(Qg)).max()
    eq2 = np.isclose(Qo, Qg, rtol=0, atol=0, equal_nan=True)
    print(f"granular (C ABI) vs reference sm_100 (-fmad=false): bit-equal {bits_equal(Qo, Qg)}, equal cells {eq2.mean():.6f}, nLinf {err2:.2e}")
    assert eq2.all()


@pytest.mark.parametrize("tag", ["A", "B"])
def test_fused_c1_100_steps_vs_reference_kernels_on_the_same_gpu(imhd, torch, O, oracle_mod, tag):
    G = refgpu(oracle_mod, nofma=False)  # stock flags
    path, D = (oracle_mod.PATH_A, 0.0) if tag == "A" else (oracle_mod.PATH_B, D_B)
    dims = (64, 64, 128)  # C1
    _, d, Q0 = make_case(O, oracle_mod, *dims)
    geom = G.COVER_A if tag == "A" else G.COVER_B
    Qr, _ = run_ref_gpu(torch, G, Q0, path, D, d, 100, geom)
    with imhd.ops.Context(*dims) as ctx:
        ctx.set_state(Q0)
        ctx.set_spacing(*d)
        ctx.prime(path, D, DT)
        ctx.step(100)
        Qf = ctx.get_state()
    assert np.isfinite(Qr).all() and np.isfinite(Qf).all()
    err = oracle_mod.normalised_linf(Qf, Qr)
    print(f"\nfused vs reference kernels on sm_100 (stock flags), C1 path {tag}, 100 steps: nLinf per variable {err}")
    assert err.max() <= TOL


@pytest.mark.parametrize("tag", ["A", "B"])
def test_production_grid_100_steps_vs_reference_kernels_on_the_same_gpu(imhd, torch, oracle_mod, tag):
    """north_star bar at the production size: 304x304x592 (BASELINE configs[1]), 100 steps, per-variable normalised
    L-inf <= 1e-5 against the reference's unmodified kernels (stock flags) run on the same B200 in the launch order of
    src/on-device/main.cu:196-213 (path B, D = 0.01) / no_diffusion.cu:284-312 (path A).  ~1 min: the reference needs
    ~0.5 s per step with diffusion at this size."""
    from conftest import BOUNDS

    G = refgpu(oracle_mod, nofma=False)
    path, D = (oracle_mod.PATH_A, 0.0) if tag == "A" else (oracle_mod.PATH_B, D_B)
    Nx, Ny, Nz = 304, 304, 592
    geom = G.COVER_A if tag == "A" else G.COVER_B
    assert G.covers(geom, Nx, Ny, Nz)
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 12 * (1 << 30):
        pytest.skip("needs ~9 GB of device memory")
    with imhd.ops.Context(Nx, Ny, Nz) as ctx:
        ctx.init_grids(*BOUNDS)
        ctx.init_screwpinch_stride(1.0)
        Q0 = ctx.get_state()
        d = tuple(float(oracle_mod.grid_spacing(BOUNDS[2 * a], BOUNDS[2 * a + 1], n)) for a, n in enumerate((Nx, Ny, Nz)))
        Q = torch.from_numpy(Q0).cuda()
        Qint = torch.zeros_like(Q)
        G.prime(Q.data_ptr(), Qint.data_ptr(), (Nx, Ny, Nz), path, D, DT, *d, geom)
        G.steps(Q.data_ptr(), Qint.data_ptr(), (Nx, Ny, Nz), path, 100, D, DT, *d, geom)
        torch.cuda.synchronize()
        del Qint
        ctx.prime(path, D, DT)
        ctx.step(100)
        Qf = torch.from_numpy(ctx.get_state()).cuda()
    assert bool(torch.isfinite(Q).all()) and bool(torch.isfinite(Qf).all())
    err = [float((Qf[v] - Q[v]).abs().max() / Q[v].abs().max()) for v in range(8)]
    print(f"\nfused vs reference kernels on sm_100 (stock flags), 304x304x592 path {tag}, 100 steps: nLinf per variable "
          + " ".join(f"{e:.2e}" for e in err))
    assert max(err) <= TOL


def test_initial_condition_kernels_vs_reference_kernels_on_the_gpu(imhd, torch, O, oracle_mod):
    """All five IC kernels of initialize_od.cu, reference (nvcc sm_100, -fmad=false) vs ours, same device: here the
    transcendental functions (logf, cosf) are the SAME libdevice code, so every one must match bit for bit."""
    from conftest import BOUNDS

    G = refgpu(oracle_mod, nofma=True)
    dims = (40, 36, 20)
    gx, gy, gz = imhd.ops.init_grids(BOUNDS, *dims)
    pre = torch.from_numpy(random_state(*dims, seed=5)).cuda()
    ours = {"screwpinch-stride": (imhd.ops.init_screwpinch_stride(1.0, gx, gy, gz), 1.0, 0.0),
            "cubic-bennett-vortex-m0": (imhd.ops.init_cubic_bennett_vortex_m0(2.0, 0.5, gx, gy, gz), 2.0, 0.5),
            "cubic-bennett-vortex": (imhd.ops.init_cubic_bennett_vortex(gx, gy, gz), 0.0, 0.0),
            "zpinch": (imhd.ops.init_zpinch(0.3, gx, gy, gz), 0.3, 0.0),
            "screwpinch": (imhd.ops.init_screwpinch(1.0, 0.3, gx, gy, gz, prefill=pre), 1.0, 0.3)}
    for key, (Q, a, b) in ours.items():
        Qr = pre.clone()
        G.init(key, Qr.data_ptr(), a, b, gx.data_ptr(), gy.data_ptr(), gz.data_ptr(), dims)
        torch.cuda.synchronize()
        eq = bits_equal(Q.cpu().numpy(), Qr.cpu().numpy())
        err = np.abs(Q.cpu().numpy().astype(np.float64) - Qr.cpu().numpy()).max()
        print(f"\n{key}: bit-identical to the reference kernel on sm_100 = {eq} (max abs diff {err:.2e})")
        assert eq, key
