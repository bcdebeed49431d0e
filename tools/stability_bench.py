"""Times the device CFL scan (imhd_stability_scan) on the bench workload and compares it with the HBM roofline
(32 B per cell read once).  MEASUREMENT TOOLING.   python tools/stability_bench.py"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
imhd = importlib.import_module("imhd-cuda_b200")
ops = imhd.ops
BOUNDS = (-3.14159, 3.14159) * 3

dims = (304, 304, 592)
gx, gy, gz = ops.init_grids(BOUNDS, *dims)
Q = ops.init_screwpinch_stride(1.0, gx, gy, gz)
d = [float((BOUNDS[2 * a + 1] - BOUNDS[2 * a]) / (n - 1)) for a, n in enumerate(dims)]
slab = ops.make_slab(*dims, 1, 0.01, 1e-4, *d)
r = ops.stability_scan(Q, slab)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    r = ops.stability_scan(Q, slab)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
cells = dims[0] * dims[1] * dims[2]
peak = 6545.3
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
gbs = cells * 32 / ms * 1e-6
print(json.dumps({"workload": "304x304x592 screw pinch", "mode": "wave speeds (closed form)", "scan": r,
                  "ms_per_scan_incl_sync_and_d2h": ms, "algorithmic_GBps": gbs, "hbm_peak_GBps": peak, "frac": gbs / peak}))
# the reference's report: closed form in x, the spectral radius of ITS y and z matrices by QR per cell
imhd._lib.load().imhd_stability_mode(1)
r = ops.stability_scan(Q, slab)
e0.record()
for _ in range(3):
    r = ops.stability_scan(Q, slab)
e1.record()
torch.cuda.synchronize()
imhd._lib.load().imhd_stability_mode(0)
print(json.dumps({"workload": "304x304x592 screw pinch", "mode": "reference-quirks (two 8x8 eigenproblems per cell, fp64 QR)",
                  "scan": r, "ms_per_scan_incl_sync_and_d2h": e0.elapsed_time(e1) / 3}))
