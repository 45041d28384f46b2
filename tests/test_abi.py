"""The C-ABI library loads and exports every symbol include/wfm_b200.h declares (no GPU needed)."""
import ctypes as C
import os

import pytest

from microtipi_b200 import _capi as capi
from tests.util import header_symbols, ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build_library()
    return capi.load_library()


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 35
    assert sorted(capi.SIGNATURES) == syms


def test_every_declared_symbol_is_exported(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_library_is_in_tree_and_reports_sm100a(lib):
    assert os.path.dirname(capi.LIB_PATH).startswith(ROOT)
    assert b"sm_100a" in lib.wfm_version()


def test_create_rejects_bad_arguments_like_the_reference(lib):
    h = C.c_void_p()
    assert lib.wfm_create(C.byref(h), 64, 32, 8, 1e-7, 1e-7, 0, 0) == capi.WFM_ERR_INVALID_ARG   # WFM:158
    assert b"Nx should equal Ny" in lib.wfm_last_error(None)
    assert lib.wfm_create(C.byref(h), 5000, 5000, 8, 1e-7, 1e-7, 0, 0) == capi.WFM_ERR_UNSUPPORTED   # any-N path: N <= 4096
    assert lib.wfm_create(C.byref(h), 64, 64, 0, 1e-7, 1e-7, 0, 0) == capi.WFM_ERR_INVALID_ARG
    assert lib.wfm_create_slab(C.byref(h), 64, 64, 8, 6, 4, 1e-7, 1e-7, 0, 0) == capi.WFM_ERR_INVALID_ARG


def test_no_cpu_fallback_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.wfm_create(C.byref(h), 64, 64, 8, 1e-7, 1e-7, 0, 0) == capi.WFM_ERR_CUDA
    assert b"no CPU fallback" in lib.wfm_last_error(None)
    assert not h.value


def test_product_package_never_imports_the_oracle():
    import pathlib
    for f in pathlib.Path(ROOT, "microtipi_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src and "tests.emu" not in src, f
