"""Shared helpers for the test-suite."""
import ctypes as C
import os
import re

import numpy as np

from oracle import wfm_oracle as o
from microtipi_b200 import _capi as capi
from microtipi_b200 import WideFieldModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = o.DEFAULTS
BETA4 = [1.0, 0.1, -0.05, 0.02]

_emu = None


def emu_lib():
    """CPU-emulated build of the SAME kernel/API sources (tests only, never the product path)."""
    global _emu
    if _emu is None:
        from tests.emu.build_emu import build
        _emu = capi.bind(C.CDLL(build()))
    return _emu


def gpu_lib():
    return capi.load_library()


def oracle_basis(N):
    return lambda nz: o.compute_zernike(nz, N, N, P["NA"], P["lam"], P["dxy"])


def make_pair(N, Nz, lib, nPhase=10, nModulus=4, single=False, device_basis=False, delta=None, **kw):
    """(oracle model, model under test) on identical synthetic pupils."""
    ref = o.WideFieldModelOracle((N, N, Nz), nPhase, nModulus, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"],
                                 single=single)
    m = WideFieldModel((N, N, Nz), nPhase, nModulus, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single,
                       lib=lib, basis=None if device_basis else oracle_basis(N), **kw)
    alpha = o.synthetic_alpha(nPhase) if nPhase else None
    beta = BETA4[:nModulus] if nModulus > 1 else [1.0]
    for mm in (ref, m):
        if nPhase:
            mm.setPhase(alpha)
        mm.setModulus(beta)
        if delta is not None:
            mm.setDefocus([P["ni"] / P["lam"], delta[0], delta[1]])
    return ref, m


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "wfm_b200.h")).read()
    return sorted(set(re.findall(r"WFM_API\s+[\w\s\*]+?\b(wfm_\w+)\s*\(", txt)))


def tol(single):
    """north_star tolerances: rel-L2 <= 1e-12 (fp64), <= 1e-5 (fp32)."""
    return 1e-5 if single else 1e-12
