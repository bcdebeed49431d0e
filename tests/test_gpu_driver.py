"""End-to-end drop-in test (GPU): the launcher + imhd-cuda / imhd-cuda_nodiff executables with the reference's
argv lists, reading back the .h5 frames they write and comparing with the CPU oracle."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import BOUNDS, ROOT, make_case
from h5_min_reader import Reader

pytestmark = pytest.mark.gpu
VARS = ("rho", "rhovx", "rhovy", "rhovz", "Bx", "By", "Bz", "e")
DRV = os.path.join(ROOT, "imhd-cuda_b200", "driver")


def read_frame(path, dims):
    Nx, Ny, Nz = dims
    r = Reader(path)
    return np.stack([r.dataset(v)[0].reshape(Nz, Nx, Ny) for v in VARS])


def patched_input(tmp_path, src, **over):
    lines = []
    for line in open(os.path.join(DRV, src)):
        k, v = line.strip().split("=", 1)
        lines.append(f"{k}={over.get(k, v)}")
    p = tmp_path / src
    p.write_text("\n".join(lines) + "\n")
    return str(p)


@pytest.mark.parametrize("mode", ["nodiff", "diffusion"])
def test_launcher_runs_the_drop_in_driver(tmp_path, O, oracle_mod, mode):
    om = oracle_mod
    dims = (32, 28, 20)
    Nx, Ny, Nz = dims
    nt = 13
    data = str(tmp_path / "data") + "/"
    inp = patched_input(tmp_path, "input.inp" if mode == "nodiff" else "input_diffusion.inp", Nt=nt, Nx=Nx, Ny=Ny, Nz=Nz)
    env = dict(os.environ, IMHD_OUTPUT_EVERY="4")
    os.makedirs(data)
    open(data + "stale.h5", "w").write("x")      # the launcher wipes the data directory first
    open(data + "README.md", "w").write("keep")  # ... except README.md
    out = subprocess.run([sys.executable, os.path.join(DRV, "simulation_launcher.py"), mode, "--input", inp, "--data-dir", data],
                         capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    files = sorted(os.listdir(data))
    assert "stale.h5" not in files and "README.md" in files and "grid.h5" in files
    frames = sorted(int(f[len("fluidvars_"):-3]) for f in files if f.startswith("fluidvars_"))
    assert frames == [0, 4, 8, 12], frames   # it % 4 == 0 and the last step Nt-1
    # grid.h5
    g = Reader(data + "grid.h5")
    go = O.init_grids(BOUNDS, *dims)
    for name, ref in zip(("x_grid", "y_grid", "z_grid"), go):
        arr, attrs = g.dataset(name)
        assert np.array_equal(arr, ref) and attrs["dimension"] == len(ref)
    # frame 0 = initial condition (+ attributes); last frame = Nt-1 steps of the reference pipeline
    ic = "bennett" if mode == "nodiff" else "screwpinch"
    path, D = (om.PATH_A, 0.0) if mode == "nodiff" else (om.PATH_B, 0.01)
    _, d, Q0 = make_case(O, om, *dims, ic=ic)
    r0 = Reader(data + "fluidvars_0.h5")
    assert r0.dataset("rho")[1]["cubeDimensionsNames"] == ["Nx", "Ny", "Nz"]
    assert Reader(data + "fluidvars_4.h5").dataset("rho")[1] == {}
    qo, io = Q0.copy(), np.zeros_like(Q0)
    om_dt = 1e-4
    O.prime(qo, io, path, D, om_dt, *d)
    f0 = read_frame(data + "fluidvars_0.h5", dims)
    np.testing.assert_allclose(f0, qo, rtol=0, atol=5e-7)   # path A: after the initial wall/PBC pass; GPU logf/cosf within 2 ulp
    O.steps(qo, io, path, nt - 1, D, om_dt, *d)
    fl = read_frame(data + f"fluidvars_{nt - 1}.h5", dims)
    assert (om.normalised_linf(fl, qo) <= 1e-5).all()


def test_driver_selects_the_initial_condition_by_registry_key(tmp_path, O, oracle_mod):
    """IMHD_IC=<configurers.hpp key>: same executable, another IC kernel (the reference needs a rebuild for that,
    no_diffusion.cu:168-171)."""
    om = oracle_mod
    dims = (32, 28, 20)
    data = str(tmp_path / "data") + "/"
    os.makedirs(data)
    inp = patched_input(tmp_path, "input.inp", Nt=5, Nx=dims[0], Ny=dims[1], Nz=dims[2], r_max_coeff=0.3)
    run = [sys.executable, os.path.join(DRV, "simulation_launcher.py"), "nodiff", "--input", inp, "--data-dir", data]
    out = subprocess.run(run, capture_output=True, text=True, env=dict(os.environ, IMHD_IC="zpinch", IMHD_OUTPUT_EVERY="4"))
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Initial condition: zpinch" in out.stdout
    g = O.init_grids(BOUNDS, *dims)
    d = tuple(float(om.grid_spacing(BOUNDS[2 * a], BOUNDS[2 * a + 1], n)) for a, n in enumerate(dims))
    qo = O.zpinch(0.3, *g)  # r_max_coeff from argv slot 7
    io = np.zeros_like(qo)
    O.prime(qo, io, om.PATH_A, 0.0, 1e-4, *d)
    assert np.array_equal(read_frame(data + "fluidvars_0.h5", dims), qo)
    O.steps(qo, io, om.PATH_A, 4, 0.0, 1e-4, *d)
    assert (om.normalised_linf(read_frame(data + "fluidvars_4.h5", dims), qo) <= 1e-5).all()
    # eigen_bin_name != none: the CFL report the reference's forked scanner prints (compute_stability.cpp:143-147)
    inp2 = patched_input(tmp_path, "input.inp", Nt=2, Nx=dims[0], Ny=dims[1], Nz=dims[2], eigen_bin_name="on-device/utils/fj_evs_compute")
    rep = subprocess.run([*run[:4], inp2, *run[5:]], capture_output=True, text=True, env=dict(os.environ, IMHD_IC="zpinch:0.3"))
    assert rep.returncode == 0, rep.stdout + rep.stderr
    import re
    from oracle import stability as st
    m = re.search(r"Largest violation: (\S+) at \(i,j,k\) = \((\d+),(\d+),(\d+)\)\nNew timestep: (\S+)\nTotal number of stability violations detected: (\d+)", rep.stdout)
    assert m, rep.stdout
    q0 = O.zpinch(0.3, *g)
    i0 = np.zeros_like(q0)
    O.prime(q0, i0, om.PATH_A, 0.0, 1e-4, *d)
    want = st.scan(st.wave_speed_lhs(q0, 1e-4, *d), 1e-4)
    assert float(m.group(1)) == pytest.approx(want["max_lhs"], rel=1e-5) and int(m.group(6)) == want["violations"]
    assert float(m.group(5)) == pytest.approx(want["dt_new"], rel=1e-5)
    bad = subprocess.run(run, capture_output=True, text=True, env=dict(os.environ, IMHD_IC="orszag-tang"))
    assert bad.returncode != 0 and "Unknown simulation type: orszag-tang" in bad.stdout + bad.stderr


def test_context_output_from_python(tmp_path, O, oracle_mod):
    """Context.write_frame / write_grid / flush_output: frames queued while the time loop keeps running hold the state
    of the step they were queued at."""
    imhd = importlib.import_module("imhd-cuda_b200")
    om = oracle_mod
    dims = (24, 20, 12)
    _, d, Q0 = make_case(O, om, *dims)
    out = str(tmp_path / "frames")
    os.makedirs(out)
    states = {}
    with imhd.Context(*dims) as c:
        c.init_grids(*BOUNDS)
        c.set_state(Q0)
        c.prime(om.PATH_B, 0.01, 1e-4)
        c.write_grid(out)
        for it in range(0, 7):
            if it:
                c.step(1)
            if it % 2 == 0:
                c.write_frame(out, it)        # no wait: the next step overwrites the device buffers
                states[it] = None
        c.flush_output()
        final = c.get_state()
    assert sorted(f for f in os.listdir(out)) == ["fluidvars_0.h5", "fluidvars_2.h5", "fluidvars_4.h5", "fluidvars_6.h5", "grid.h5"]
    assert np.array_equal(read_frame(os.path.join(out, "fluidvars_6.h5"), dims), final)
    qo, io = Q0.copy(), np.zeros_like(Q0)
    O.prime(qo, io, om.PATH_B, 0.01, 1e-4, *d)
    assert np.array_equal(read_frame(os.path.join(out, "fluidvars_0.h5"), dims), qo)
    O.steps(qo, io, om.PATH_B, 2, 0.01, 1e-4, *d)
    assert (om.normalised_linf(read_frame(os.path.join(out, "fluidvars_2.h5"), dims), qo) <= 1e-6).all()
