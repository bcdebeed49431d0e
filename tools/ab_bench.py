"""A/B timing of fused-kernel variants on the bench workload (304x304x592, screw pinch), one GPU.

    python tools/ab_bench.py [--steps 30] [--paths B,A] [--variants 0,16,48] [--check]

`variant` is the flag word of imhd_set_kernel_variant (bits 4..7 choose the kernel: 0 default, 1 one row per
thread, 3/4/5 other register tilings).  Times K steps with CUDA events on the context's stream after warm-up and
prints ms/step and GLUPS per (path, variant); --check also compares the final states bit for bit against the
first variant in the list.  Measurement helper, not part of the library.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--paths", default="B")
    ap.add_argument("--variants", default="0,16")
    ap.add_argument("--dims", default="304,304,592")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch

    imhd = importlib.import_module("imhd-cuda_b200")
    lib = imhd._lib.load()
    Nx, Ny, Nz = (int(x) for x in args.dims.split(","))
    B = (-3.14159, 3.14159) * 3
    rows = []
    for path_name in args.paths.split(","):
        path, D = (imhd.PATH_B, 0.01) if path_name == "B" else (imhd.PATH_A, 0.0)
        ref = None
        for var in (int(v) for v in args.variants.split(",")):
            lib.imhd_set_kernel_variant(var)
            with imhd.Context(Nx, Ny, Nz) as c:
                c.init_grids(*B)
                c.init_screwpinch_stride(1.0)
                c.prime(path, D, 1e-4)
                c.step(args.warmup)
                c.synchronize()
                st = torch.cuda.ExternalStream(lib.imhd_ctx_stream(c.h)) if lib.imhd_ctx_stream(c.h) else torch.cuda.current_stream()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                c.step(args.steps)
                e1.record(st)
                c.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                row = {"path": path_name, "variant": var, "ms_per_step": round(ms, 4), "glups": round(Nx * Ny * Nz / ms / 1e6, 2)}
                if args.check:
                    Q = c.get_state()
                    row["finite"] = bool(np.isfinite(Q).all())
                    if ref is None:
                        ref = Q
                    else:
                        row["bit_identical_to_first"] = bool(np.array_equal(ref.view(np.uint32), Q.view(np.uint32)))
                rows.append(row)
                print(json.dumps(row), flush=True)
    lib.imhd_set_kernel_variant(0)
    return rows


if __name__ == "__main__":
    main()
