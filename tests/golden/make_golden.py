"""Generate the committed golden vectors from the oracle (run in the build container):

    python tests/golden/make_golden.py

The reference itself ships no fixtures and cannot be run here (no JDK / TiPi / JTransforms), so
these vectors freeze the oracle's output -- itself pinned by the known-answer tests of
tests/test_oracle.py -- on the SURVEY.md 8d2 synthetic inputs.  The GPU parity tests compare the
CUDA path with them; tests/test_golden.py checks that the oracle still reproduces them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import wfm_oracle as o  # noqa: E402

P = o.DEFAULTS
BETA4 = [1.0, 0.1, -0.05, 0.02]
CPX_PLANES_CFG1 = [0, 1, 16, 17, 31]


def case(N, Nz, single, delta=None):
    m = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], single=single)
    m.setPhase(o.synthetic_alpha(10))
    m.setModulus(BETA4)
    if delta is not None:
        m.setDefocus([P["ni"] / P["lam"], delta[0], delta[1]])
    q = o.synthetic_q(N, N, Nz, single=single)
    out = dict(N=N, Nz=Nz, single=single, delta=np.array(delta if delta else [0.0, 0.0]),
               alpha=m.alpha, beta=m.beta, psf=m.getPsf(),
               j_phase=m.apply_J_phase(q), j_defocus=m.apply_J_defocus(q), j_modulus=m.apply_J_modulus(q))
    m.modulus_mode = o.MODULUS_REFERENCE_LAST_PLANE
    out["j_modulus_last_plane"] = m.apply_J_modulus(q)
    return m, out


def main():
    m, out = case(32, 8, False)
    out["cpx"] = m.get_cpxPsf()
    np.savez_compressed(os.path.join(HERE, "wfm_n32_z8_f64.npz"), **out)
    m, out = case(32, 8, True)
    out["cpx"] = m.get_cpxPsf()
    np.savez_compressed(os.path.join(HERE, "wfm_n32_z8_f32.npz"), **out)
    m, out = case(32, 8, False, delta=(2e4, -2e4))
    out["cpx"] = m.get_cpxPsf()
    np.savez_compressed(os.path.join(HERE, "wfm_n32_z8_f64_offaxis.npz"), **out)
    # BASELINE config 1: 64x64x32, NA 1.4, 10 Zernike phase coefficients, fp64
    m, out = case(64, 32, False)
    out["cpx_planes"] = np.array(CPX_PLANES_CFG1)
    out["cpx"] = m.get_cpxPsf()[CPX_PLANES_CFG1]
    np.savez_compressed(os.path.join(HERE, "wfm_n64_z32_f64_config1.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
