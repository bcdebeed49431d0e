"""Profiling workload: a few fused steps on the production grid (used under ncu; never a bench value)."""
import argparse, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--path", type=int, default=1)
ap.add_argument("--dims", type=int, nargs=3, default=[304, 304, 592])
a = ap.parse_args()
pkg = importlib.import_module("imhd-cuda_b200")
Nx, Ny, Nz = a.dims
with pkg.Context(Nx, Ny, Nz) as c:
    c.init_grids(*((-3.14159, 3.14159) * 3))
    c.init_screwpinch_stride(1.0)
    c.prime(a.path, 0.01 if a.path else 0.0, 1e-4)
    c.step(a.steps)
    c.synchronize()
print("done", pkg.ops.launch_count())
