"""The .h5 output contract written without libhdf5, read back with the independent reader
in tests/h5_min_reader.py.  Host-only code: runs without a GPU."""
import ctypes as C

import numpy as np

from h5_min_reader import Reader

VARS = ("rho", "rhovx", "rhovy", "rhovz", "Bx", "By", "Bz", "e")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_fluidvars_frame_with_attributes(imhd, tmp_path):
    lib = imhd._lib.load()
    Nx, Ny, Nz = 6, 5, 7
    Q = np.arange(8 * Nx * Ny * Nz, dtype=np.float32).reshape(8, Nz, Nx, Ny) * 0.5
    path = str(tmp_path / "fluidvars_0.h5")
    assert lib.imhd_h5_write_fluidvars(path.encode(), _p(Q), Nx, Ny, Nz, 1) == 0
    r = Reader(path)
    assert r.names() == sorted(VARS)
    for v, name in enumerate(VARS):
        data, attrs = r.dataset(name)
        assert data.dtype == np.float32 and data.shape == (Nx * Ny * Nz,)          # 1-D, IDX3D order (phdf5_write_all.cpp:96-97)
        assert np.array_equal(data, Q[v].ravel())
        assert list(attrs["cubeDimensions"]) == [Nx, 0, Ny]                        # the reference's hsize_t-as-int quirk (B-21)
        assert attrs["cubeDimensionsNames"] == ["Nx", "Ny", "Nz"]
        assert attrs["storagePattern"] == ["Row-major, depth-minor: l = k * (Nx * Ny) + i * Ny + j"]


def test_later_frames_carry_no_attributes(imhd, tmp_path):
    lib = imhd._lib.load()
    Q = np.random.default_rng(1).standard_normal((8, 4, 4, 4)).astype(np.float32)
    path = str(tmp_path / "fluidvars_7.h5")
    assert lib.imhd_h5_write_fluidvars(path.encode(), _p(Q), 4, 4, 4, 0) == 0
    r = Reader(path)
    for v, name in enumerate(VARS):
        data, attrs = r.dataset(name)
        assert np.array_equal(data, Q[v].ravel()) and attrs == {}


def test_grid_file(imhd, tmp_path):
    lib = imhd._lib.load()
    x = np.linspace(-3.14159, 3.14159, 9, dtype=np.float32)
    y = np.linspace(-3.14159, 3.14159, 11, dtype=np.float32)
    z = np.linspace(-3.14159, 3.14159, 5, dtype=np.float32)
    path = str(tmp_path / "grid.h5")
    assert lib.imhd_h5_write_grid(path.encode(), _p(x), _p(y), _p(z), 9, 11, 5) == 0
    r = Reader(path)
    assert r.names() == ["x_grid", "y_grid", "z_grid"]
    for name, g in (("x_grid", x), ("y_grid", y), ("z_grid", z)):
        data, attrs = r.dataset(name)
        assert np.array_equal(data, g)
        assert attrs["dimension"] == len(g) and attrs["dimension"].dtype == np.int32
        assert attrs["spacing"] == np.float32((g[-1] - g[0]) / np.float32(len(g) - 1))   # hdf5_write_grid.cpp:91-93


def test_unwritable_path_is_an_error(imhd):
    lib = imhd._lib.load()
    Q = np.zeros((8, 4, 4, 4), np.float32)
    assert lib.imhd_h5_write_fluidvars(b"/nonexistent-dir/x.h5", _p(Q), 4, 4, 4, 0) == imhd._lib.E_IO
    assert b"cannot open" in lib.imhd_last_error()
