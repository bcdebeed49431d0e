"""Where does the strip variant differ from the no-strip one?  python tools/experiments/strip_diff.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
imhd = importlib.import_module("imhd-cuda_b200")
from oracle import oracle as om
from test_oracle_golden import random_state
lib = imhd._lib.load(); O = om.Oracle(); B = (-3.14159, 3.14159) * 3
for dims in ((40, 64, 21), (64, 64, 40)):
    g = O.init_grids(B, *dims)
    d = tuple(float(om.grid_spacing(B[2 * a], B[2 * a + 1], n)) for a, n in enumerate(dims))
    Q0 = O.cubic_bennett_vortex_m0(2.0, 0.5, *g) + 0.01 * random_state(*dims, seed=3)
    for path, D in ((0, 0.0), (1, 0.01)):
        out = []
        for var in (2, 4, 4 | 256):
            lib.imhd_set_kernel_variant(var)
            with imhd.Context(*dims) as c:
                c.set_state(Q0); c.set_spacing(*d); c.prime(path, D, 1e-4); c.step(1); out.append(c.get_state())
        lib.imhd_set_kernel_variant(0)
        for name, o in (("wstrip", out[1]), ("blockstrip", out[2])):
            diff = out[0].view(np.uint32) != o.view(np.uint32)
            print(dims, "path", path, name, "differing values:", int(diff.sum()))
            if diff.any():
                v, k, i, j = np.nonzero(diff)
                print("   v:", sorted(set(v.tolist())), " k:", sorted(set(k.tolist()))[:12], " i:", sorted(set(i.tolist()))[:20], " j:", sorted(set(j.tolist())))
                idx = (v[0], k[0], i[0], j[0]); print("   first", idx, out[0][idx], o[idx])
