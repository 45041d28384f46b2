"""The header-only C++ host mirror (include/wfm_b200.hpp) builds and behaves like the reference class."""
import os
import subprocess

import pytest

from tests.util import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_wfm_hpp.cpp")


def _build_and_run(libdir, libname, tmp_path):
    exe = str(tmp_path / "test_wfm_hpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", libdir, f"-l:{libname}", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK")


def test_cpp_mirror_on_emulated_library(tmp_path):
    from tests.emu.build_emu import build
    lib = build()
    _build_and_run(os.path.dirname(lib), os.path.basename(lib), tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path):
    import __graft_entry__ as g
    lib = g.build_library()
    _build_and_run(os.path.dirname(lib), os.path.basename(lib), tmp_path)
