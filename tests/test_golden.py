"""The oracle reproduces the committed golden vectors (guards the checker against drift)."""
import glob
import os

import numpy as np
import pytest

from oracle import wfm_oracle as o
from tests.golden.make_golden import case

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "*.npz"))))
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    delta = tuple(g["delta"]) if np.any(g["delta"]) else None
    m, out = case(int(g["N"]), int(g["Nz"]), bool(g["single"]), delta)
    t = 1e-6 if bool(g["single"]) else 1e-13
    assert o.rel_l2(out["psf"], g["psf"]) <= t
    for k in ("j_phase", "j_defocus", "j_modulus", "j_modulus_last_plane"):
        assert o.rel_l2(out[k], g[k]) <= 50 * t, k
    cpx = m.get_cpxPsf()
    if "cpx_planes" in g:
        cpx = cpx[g["cpx_planes"]]
    assert o.rel_l2(cpx, g["cpx"]) <= t
