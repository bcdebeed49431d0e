"""Times the reference's OWN kernels, recompiled by nvcc for sm_100 (oracle/_ref/libimhd_ref_gpu.so, stock flags),
on the B200 -- SURVEY.md 8(d) "reference GPU number on the same box": the kernels to beat -- next to the fused
step of libimhd_b200.so on the same grid and initial state.  MEASUREMENT TOOLING, not product code.

    python tools/ref_gpu_bench.py [--out gpurun_out/ref_gpu_bench.json] [--big-steps 3]

Launch geometries (see oracle/ref_shim/ref_gpu_harness.cu):
  stock     the shipped input.inp values (6x6x6 blocks, grid = 3 x numberOfSMs per axis)    [path A only]
  cover     8x8x4 blocks, smallest grid covering the domain
  coalesced 1x32x8 blocks (warp lanes along the unit-stride axis), smallest covering grid
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as om  # noqa: E402

imhd = importlib.import_module("imhd-cuda_b200")
BOUNDS = (-3.14159, 3.14159) * 3
DT, D_B = 1e-4, 0.01


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu_bench.json"))
    ap.add_argument("--big-steps", type=int, default=3)
    a = ap.parse_args()
    G = om.ReferenceGPU(nofma=False)
    rows = []
    for name, dims, nsteps in (("C1 64x64x128", (64, 64, 128), 20), ("C2 304x304x592", (304, 304, 592), a.big_steps)):
        Nx, Ny, Nz = dims
        cells = Nx * Ny * Nz
        d = tuple(float(om.grid_spacing(BOUNDS[2 * q], BOUNDS[2 * q + 1], n)) for q, n in enumerate(dims))
        gx, gy, gz = imhd.ops.init_grids(BOUNDS, *dims)
        Q0 = imhd.ops.init_screwpinch_stride(1.0, gx, gy, gz)
        for tag, path, D in (("A", om.PATH_A, 0.0), ("B", om.PATH_B, D_B)):
            geoms = [("cover", G.COVER_A if tag == "A" else G.COVER_B),
                     ("coalesced", G.COALESCED_A if tag == "A" else G.COALESCED_B)]
            if tag == "A":
                geoms.insert(0, ("stock", G.STOCK_A))
            for gname, geom in geoms:
                if not G.covers(geom, *dims):
                    continue
                Q = Q0.clone()
                Qint = torch.zeros_like(Q)
                G.prime(Q.data_ptr(), Qint.data_ptr(), dims, path, D, DT, *d, geom)
                G.steps(Q.data_ptr(), Qint.data_ptr(), dims, path, 1, D, DT, *d, geom)  # warm-up
                ms = G.steps(Q.data_ptr(), Qint.data_ptr(), dims, path, nsteps, D, DT, *d, geom)
                per = [m / nsteps for m in ms]
                row = {"impl": "reference kernels, nvcc sm_100, stock flags", "config": name, "path": tag,
                       "geometry": gname, "geom": list(geom), "steps": nsteps, "ms_per_step": per[4],
                       "ms_corrector": per[0], "ms_fluid_bcs": per[1], "ms_predictor": per[2],
                       "ms_qint_boundary": per[3], "glups": cells / per[4] * 1e-6,
                       "finite": bool(torch.isfinite(Q).all().item())}
                rows.append(row)
                print(json.dumps(row), flush=True)
                del Q, Qint
            # ours, same grid / IC, fused step through the C ABI
            with imhd.ops.Context(*dims) as ctx:
                ctx.set_state(Q0.cpu().numpy())
                ctx.set_spacing(*d)
                ctx.prime(path, D, DT)
                ctx.step(5)
                ctx.synchronize()
                n = 50
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s = torch.cuda.ExternalStream(ctx.L.imhd_ctx_stream(ctx.h))
                e0.record(s)
                ctx.step(n)
                e1.record(s)
                ctx.synchronize()
                t = e0.elapsed_time(e1) / n
            row = {"impl": "libimhd_b200 fused step", "config": name, "path": tag, "steps": n, "ms_per_step": t,
                   "glups": cells / t * 1e-6}
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
