"""Where does the no-diffusion pipeline (path A) leave the finite numbers on the production grid?  Ours vs the
reference's own kernels on the same GPU, screw pinch and Bennett vortex.  Measurement helper."""
import importlib, math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import oracle as om
imhd = importlib.import_module("imhd-cuda_b200")
B = (-3.14159, 3.14159) * 3
Nx, Ny, Nz = 304, 304, 592
d = tuple(float(om.grid_spacing(B[2 * a], B[2 * a + 1], n)) for a, n in enumerate((Nx, Ny, Nz)))
G = om.ReferenceGPU(nofma=False)
for ic in ("screwpinch-stride", "cubic-bennett-vortex-m0"):
    params = (1.0,) if ic.startswith("screw") else (2 * math.pi * 2 / (B[5] - B[4]), 0.5)
    with imhd.Context(Nx, Ny, Nz) as c:
        c.init_grids(*B); c.initialize(ic, *params)
        Q0 = c.get_state()
        c.prime(imhd.PATH_A, 0.0, 1e-4)
        first = None
        for n in range(0, 400, 25):
            c.step(25)
            st = torch.as_tensor(type("D", (), {"__cuda_array_interface__": {"shape": (8, Nz, Nx, Ny), "typestr": "<f4", "data": (int(imhd._lib.load().imhd_ctx_device_state(c.h)), False), "version": 3}})(), device="cuda")
            c.synchronize()
            if not bool(torch.isfinite(st).all()):
                first = n + 25; break
        print(ic, "ours: first non-finite state within step", first, flush=True)
    Q = torch.from_numpy(Q0).cuda(); Qi = torch.zeros_like(Q)
    G.prime(Q.data_ptr(), Qi.data_ptr(), (Nx, Ny, Nz), om.PATH_A, 0.0, 1e-4, *d, G.COVER_A)
    first = None
    for n in range(0, 400, 25):
        G.steps(Q.data_ptr(), Qi.data_ptr(), (Nx, Ny, Nz), om.PATH_A, 25, 0.0, 1e-4, *d, G.COVER_A)
        torch.cuda.synchronize()
        if not bool(torch.isfinite(Q).all()):
            first = n + 25; break
    print(ic, "reference kernels (sm_100): first non-finite state within step", first, flush=True)
    del Q, Qi
