"""Multi-GPU engine through the C ABI (imhd_create_multi): the z-slab time loop in C++ over real NCCL/NVLink must
give the single-GPU bits for any number of slabs (SURVEY.md 8e "Parity across P").  Needs >= 2 GPUs on the box
(`gpurun --gpus 2|4 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a one-GPU box."""
import numpy as np
import pytest

from conftest import BOUNDS, bits_equal, make_case

pytestmark = pytest.mark.gpu
DT, D_B = 1e-4, 0.01


@pytest.fixture(scope="module")
def ngpu():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("multi-GPU engine needs >= 2 GPUs on the box")
    return n


def single(imhd, dims, setup, path, D, nsteps):
    with imhd.ops.Context(*dims) as c:
        setup(c)
        c.prime(path, D, DT)
        c.step(nsteps)
        return c.get_state()


@pytest.mark.parametrize("nz", [26, 61])   # 61: uneven slabs, long enough for the overlapped (ends-first) schedule at 2 slabs
@pytest.mark.parametrize("tag", ["A", "B"])
def test_multi_context_is_bit_identical_to_single_gpu(imhd, oracle_mod, O, ngpu, tag, nz):
    om = oracle_mod
    path, D = (om.PATH_A, 0.0) if tag == "A" else (om.PATH_B, D_B)
    dims = (36, 32, nz)
    _, d, Q0 = make_case(O, om, *dims, ic="bennett")
    Q0 = Q0 + 0.01 * np.random.default_rng(5).standard_normal(Q0.shape).astype(np.float32)

    def setup(c):
        c.set_state(Q0)
        c.set_spacing(*d)

    ref = single(imhd, dims, setup, path, D, 6)
    for world in sorted({2, min(ngpu, 4), min(ngpu, nz // 3, 8)}):
        with imhd.ops.Context.multi(*dims, world) as c:
            assert c.num_slabs == world
            setup(c)
            c.prime(path, D, DT)
            c.step(6)
            out = c.get_state()
            st = c.stability(DT)
        assert bits_equal(out, ref), f"path {tag}, {world} slabs over NCCL"
        with imhd.ops.Context(*dims) as s1:
            s1.set_state(ref); s1.set_spacing(*d)
            assert st == s1.stability(DT)


def test_multi_context_initial_conditions_and_run_host(imhd, oracle_mod, O, ngpu):
    """IC kernels on slab arrays (global plane offset) and the whole-job entry point give the single-GPU bits."""
    om = oracle_mod
    dims = (40, 32, 30)
    for key, params in (("screwpinch-stride", (1.0,)), ("cubic-bennett-vortex-m0", (2.0, 0.5)), ("zpinch", (0.3,))):
        def setup(c, key=key, params=params):
            c.init_grids(*BOUNDS)
            c.initialize(key, *params)

        ref = single(imhd, dims, setup, om.PATH_B, D_B, 3)
        with imhd.ops.Context.multi(*dims, 2) as c:
            setup(c)
            c.prime(om.PATH_B, D_B, DT)
            c.step(3)
            assert bits_equal(c.get_state(), ref), key
    _, d, Q0 = make_case(O, om, *dims)
    out1, out2 = np.empty_like(Q0), np.empty_like(Q0)
    with imhd.ops.Context(*dims) as c:
        c.run_host(Q0, out1, om.PATH_A, 0.0, DT, *d, 5)
    with imhd.ops.Context.multi(*dims, 2) as c:
        c.run_host(Q0, out2, om.PATH_A, 0.0, DT, *d, 5)
        k0, nzl, dev = c.slab_extent(1)
        assert (k0, nzl, dev) == (15, 15, 1)
        assert bits_equal(c.get_state_local(1), out2[:, k0:k0 + nzl])
    assert bits_equal(out1, out2)


def test_driver_with_two_gpus_writes_the_same_files(tmp_path, ngpu):
    """bin/imhd-cuda with IMHD_GPUS=2 (imhd_create_multi behind the reference's argv surface): byte-identical .h5 frames."""
    import filecmp
    import os
    import subprocess
    import sys

    from conftest import ROOT

    drv = os.path.join(ROOT, "imhd-cuda_b200", "driver")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_driver import patched_input

    outs = []
    for gpus in ("1", "2"):
        data = str(tmp_path / f"data{gpus}") + "/"
        os.makedirs(data)
        inp = patched_input(tmp_path, "input_diffusion.inp", Nt=9, Nx=32, Ny=28, Nz=24)
        env = dict(os.environ, IMHD_OUTPUT_EVERY="4", IMHD_GPUS=gpus)
        out = subprocess.run([sys.executable, os.path.join(drv, "simulation_launcher.py"), "diffusion", "--input", inp, "--data-dir", data],
                             capture_output=True, text=True, env=env)
        assert out.returncode == 0, out.stdout + out.stderr
        outs.append(data)
    files = sorted(os.listdir(outs[0]))
    assert files == sorted(os.listdir(outs[1])) and "fluidvars_8.h5" in files and "grid.h5" in files
    for f in files:
        assert filecmp.cmp(outs[0] + f, outs[1] + f, shallow=False), f


def test_slab_engine_one_process_per_slab(ngpu):
    """imhd_create_slab, the engine bench.py runs at N > 1: one process per GPU under torchrun, planes exchanged by the copy
    engines through CUDA IPC mappings (and, second run, over ncclSend / ncclRecv): the single-GPU bits, both pipelines."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = min(ngpu, 4)
    for mode in ("direct", "nccl"):
        env = dict(os.environ, IMHD_SLAB_EXCHANGE=mode)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                            "--master-port", "29533", os.path.join(root, "tools", "check_slab_engine.py")],
                           capture_output=True, text=True, timeout=300, env=env, cwd=root)
        assert r.returncode == 0, (mode, r.stdout[-2000:], r.stderr[-2000:])
        assert r.stdout.count("bit-identical = True") == 6, (mode, r.stdout)
