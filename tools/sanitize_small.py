"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
PSF + three Jacobians (narrow and generic kernels, fp64 and fp32), data term, eval_fg, rolled PSF, MTF."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microtipi_b200 import WideFieldModel, WeightedConvolutionCost, DoubleShapedVectorSpace  # noqa: E402

P = dict(NA=1.4, lam=542e-9, ni=1.518, dxy=64.5e-9, dz=160e-9)
sizes = [(64, 32)] + ([(512, 4)] if "--big" in sys.argv else [])
for N, Nz in sizes:
    for single in (False, True):
        for narrow in (True, False):
            if narrow:
                os.environ.pop("WFM_NO_NARROW", None)
            else:
                os.environ["WFM_NO_NARROW"] = "1"
            m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single)
            m.setPhase(np.linspace(-0.3, 0.3, 10))
            m.setModulus([1.0, 0.1, -0.05, 0.02])
            psf = m.getPsf()
            q = np.random.default_rng(0).uniform(-1, 1, psf.shape).astype(psf.dtype)
            d, p, mo = m.apply_J_all(q)
            assert np.all(np.isfinite(d)) and np.all(np.isfinite(p)) and np.all(np.isfinite(mo))
            assert abs(float(psf.sum(dtype=np.float64)) - 1.0) < (1e-4 if single else 1e-12)
            if not single and Nz >= 32:
                f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, Nz))
                obj = np.zeros((Nz, N, N)); obj[0, 0, 0] = 1.0; obj[0, 0, 1] = 0.5
                f.setPSF(obj); f.setData(0.9 * psf); f.setWeights(np.ones_like(psf))
                g = np.zeros(psf.size)
                c = f.computeCostAndGradient(1.0, psf, g, True)
                c2, gx = f.evalFG(m, m.PHASE, np.linspace(-0.2, 0.2, 10))
                assert np.isfinite(c) and np.isfinite(c2) and np.all(np.isfinite(gx))
                r = m.getPsfRolled(); mt = m.getMtf()
                assert np.isfinite(r).all() and np.isfinite(mt).all()
                f.close()
            m.close()
print("sanitize_small: OK")
