"""A/B helper for kernel experiments: runs a few fused steps with imhd_set_kernel_variant(0) and with the flag given
on the command line and reports whether the states are bit-identical.   python tools/experiments/variant_check.py 8"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
imhd = importlib.import_module("imhd-cuda_b200")
from oracle import oracle as om  # noqa: E402
from test_oracle_golden import random_state  # noqa: E402

flag = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = imhd._lib.load()
O = om.Oracle()
B = (-3.14159, 3.14159) * 3
ok = True
for dims in ((52, 44, 33), (40, 64, 70), (64, 96, 24)):
    g = O.init_grids(B, *dims)
    d = tuple(float(om.grid_spacing(B[2 * a], B[2 * a + 1], n)) for a, n in enumerate(dims))
    Q0 = O.cubic_bennett_vortex_m0(2.0, 0.5, *g) + 0.01 * random_state(*dims, seed=3)
    for path, D in ((0, 0.0), (1, 0.01)):
        out = []
        for var in (0, flag):
            lib.imhd_set_kernel_variant(var)
            with imhd.Context(*dims) as c:
                c.set_state(Q0); c.set_spacing(*d); c.prime(path, D, 1e-4); c.step(7); out.append(c.get_state())
        lib.imhd_set_kernel_variant(0)
        same = np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))
        ok &= same
        print(dims, "path", path, f"variant {flag} bit-identical:", same, "" if same else f"max abs diff {np.nanmax(np.abs(out[0] - out[1])):.3e}")
sys.exit(0 if ok else 1)
