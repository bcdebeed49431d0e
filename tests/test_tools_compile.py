"""Every script under tools/ and the repo-root entry points at least byte-compile (they run only on the GPU box)."""
import glob
import os
import py_compile

from conftest import ROOT


def test_scripts_byte_compile(tmp_path):
    files = sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + [os.path.join(ROOT, f) for f in ("bench.py", "__graft_entry__.py")]
    assert len(files) >= 6
    for n, f in enumerate(files):
        py_compile.compile(f, cfile=str(tmp_path / f"{n}.pyc"), doraise=True)
