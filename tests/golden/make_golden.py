"""Generate the committed golden fixtures from the REFERENCE ITSELF.

Run in the build container (needs /root/reference and oracle/_ref built):

    python tests/golden/make_golden.py

Sources of truth
  * oracle/_ref/libimhd_ref_cpu.so -- the reference's unmodified lib/on-device/*.cu compiled for the
    host (oracle/Makefile `ref`), driven in the launch order of src/on-device/no_diffusion.cu (path A)
    and src/on-device/main.cu (path B); see oracle/ref_shim/ref_harness.cpp for the launch geometry.
  * /root/reference/debug/data/rhovz/var_0.csv -- the only golden vector the reference ships
    (rho*v_z of ScrewPinchStride at t=0, 64^3, domain +-3.14159, J0=1).

Outputs (all small, committed):
  small_<ic>_path<A|B>.npz   full fp32 states of a 16x12x10 grid: primed Q/Qint and Q after 1, 2, 10 steps
  c1_path<A|B>.npz           64x64x128 screw pinch after 100 steps: every-4th-point sample, per-variable
                             min/max/fp64 sum, sha256 of the full fp32 array (C1 of BASELINE.json)
  ref_var0_rhovz_plane.csv   plane k=0 of the reference's var_0.csv (the file is z-invariant; checked here)
  manifest.json              parameters of every case
"""
import csv
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as om  # noqa: E402

BOUNDS = (-3.14159, 3.14159) * 3
DT = 1e-4
D_B = 0.01
SMALL = (16, 12, 10)  # Nx, Ny, Nz -- all different, to catch index mix-ups
C1 = (64, 64, 128)


def case(R, dims, ic):
    Nx, Ny, Nz = dims
    g = R.init_grids(BOUNDS, Nx, Ny, Nz)
    d = tuple(float(om.grid_spacing(BOUNDS[2 * a], BOUNDS[2 * a + 1], n)) for a, n in enumerate(dims))
    Q = R.screwpinch_stride(1.0, *g) if ic == "screwpinch" else R.cubic_bennett_vortex_m0(2.0, 0.5, *g)
    return g, d, Q


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    om.build(ref=True)
    R = om.Reference()
    manifest = {"bounds": BOUNDS, "dt": DT, "D_pathB": D_B, "J0": 1.0, "bennett": {"k": 2.0, "A": 0.5},
                "generator": "oracle/_ref (reference kernels on host), tests/golden/make_golden.py", "cases": {}}

    for ic in ("screwpinch", "bennett"):
        for path, tag in ((om.PATH_A, "A"), (om.PATH_B, "B")):
            D = D_B if path == om.PATH_B else 0.0
            g, (dx, dy, dz), Q0 = case(R, SMALL, ic)
            Q, Qi = Q0.copy(), np.zeros_like(Q0)
            R.prime(Q, Qi, path, D, DT, dx, dy, dz)
            out = {"Q_ic": Q0, "Q_primed": Q.copy(), "Qint_primed": Qi.copy()}
            done = 0
            for n in (1, 2, 10):
                R.steps(Q, Qi, path, n - done, D, DT, dx, dy, dz)
                done = n
                out[f"Q_step{n}"] = Q.copy()
                out[f"Qint_step{n}"] = Qi.copy()
            name = f"small_{ic}_path{tag}.npz"
            np.savez_compressed(os.path.join(HERE, name), **out)
            manifest["cases"][name] = {"dims": SMALL, "ic": ic, "path": tag, "D": D, "dx": dx, "dy": dy, "dz": dz}

    for path, tag in ((om.PATH_A, "A"), (om.PATH_B, "B")):
        D = D_B if path == om.PATH_B else 0.0
        g, (dx, dy, dz), Q0 = case(R, C1, "screwpinch")
        Q, Qi = Q0.copy(), np.zeros_like(Q0)
        R.prime(Q, Qi, path, D, DT, dx, dy, dz)
        R.steps(Q, Qi, path, 100, D, DT, dx, dy, dz)
        name = f"c1_path{tag}.npz"
        np.savez_compressed(os.path.join(HERE, name), sample=Q[:, ::4, ::4, ::4].copy(),
                            vmin=Q.reshape(8, -1).min(1), vmax=Q.reshape(8, -1).max(1),
                            vsum=Q.reshape(8, -1).astype(np.float64).sum(1))
        manifest["cases"][name] = {"dims": C1, "ic": "screwpinch", "path": tag, "D": D, "steps": 100,
                                   "dx": dx, "dy": dy, "dz": dz, "sha256_Q": sha(Q), "sha256_Q_ic": sha(Q0)}

    # the reference's own golden vector
    src = "/root/reference/debug/data/rhovz/var_0.csv"
    vals = np.zeros((64, 64, 64), np.float64)  # [k][i][j]
    with open(src) as f:
        rd = csv.reader(f)
        next(rd)
        for n, (v, i, j, k) in enumerate(rd):
            i, j, k = int(i), int(j), int(k)
            assert n == k * 64 * 64 + i * 64 + j, "row order is not IDX3D"
            vals[k, i, j] = float(v)
    assert all(np.array_equal(vals[0], vals[k]) for k in range(64)), "var_0.csv is not z-invariant"
    with open(os.path.join(HERE, "ref_var0_rhovz_plane.csv"), "w") as f:
        f.write("# plane k=0 of russellmatt66/imhd-CUDA debug/data/rhovz/var_0.csv (z-invariant), rows i, cols j\n")
        for i in range(64):
            f.write(",".join(repr(float(x)) for x in vals[0, i]) + "\n")
    manifest["ref_var0"] = {"dims": (64, 64, 64), "J0": 1.0, "nonzero_per_plane": int((vals[0] != 0).sum())}

    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
