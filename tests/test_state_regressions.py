"""Regression tests of the host-side state machine (C ABI + mirrors), run on the CPU-emulated build of the
same sources: basis replacement after coefficients were set, setPhase(None), the mirror state after evalFG,
argument-size checks and the stream-ordering entry points.  (Round-1 advisor findings.)"""
import ctypes as C

import numpy as np
import pytest

from oracle import wfm_oracle as o
from microtipi_b200 import WideFieldModel, WeightedConvolutionCost, DoubleShapedVectorSpace
from tests.util import BETA4, P, emu_lib, make_pair, oracle_basis


@pytest.fixture(scope="module")
def lib():
    return emu_lib()


def test_shrinking_the_basis_drops_coefficients_that_no_longer_fit(lib):
    """build_basis(13); set_phase(10); build_basis(4): the Jacobian must not index basis rows beyond nzern."""
    N, Nz = 32, 3
    ref, m = make_pair(N, Nz, lib)
    h = m.handle
    q = np.ascontiguousarray(o.synthetic_q(N, N, Nz))
    Z4 = np.ascontiguousarray(o.compute_zernike(4, N, N, P["NA"], P["lam"], P["dxy"]))
    assert lib.wfm_set_basis(h, Z4.ctypes.data_as(C.c_void_p), 4, 0) == 0
    nph, nmo = C.c_int(), C.c_int()
    assert lib.wfm_get_info(h, None, None, None, None, None, None, None, C.byref(nph), C.byref(nmo)) == 0
    assert nph.value == 0                       # 10 + 3 > 4: the phase vector is gone (setNPhase semantics)
    assert nmo.value == 4                       # 4 modulus coefficients still fit
    out = np.zeros(10)
    rc = lib.wfm_apply_j_phase(h, q.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 10)
    assert rc != 0                              # refused instead of reading garbage rows
    phi = np.empty(N * N)
    assert lib.wfm_get_phi(h, phi.ctypes.data_as(C.c_void_p)) == 0
    assert not phi.any()
    # a basis too small for the modulus vector drops that as well: the PSF then needs a new setModulus
    Z2 = np.ascontiguousarray(Z4[:2])
    assert lib.wfm_set_basis(h, Z2.ctypes.data_as(C.c_void_p), 2, 0) == 0
    assert lib.wfm_compute_psf(h) != 0
    beta = np.array([1.0, 0.1])
    assert lib.wfm_set_modulus(h, beta.ctypes.data_as(C.c_void_p), 2) == 0
    assert lib.wfm_compute_psf(h) == 0
    m.close()


def test_set_phase_none_reaches_the_device(lib):
    N, Nz = 32, 4
    ref, m = make_pair(N, Nz, lib)
    psf_with = m.getPsf().copy()
    m.setPhase(None)
    assert m.getNPhase() == 0 and m.PState == 0
    nph = C.c_int()
    assert lib.wfm_get_info(m.handle, None, None, None, None, None, None, None, C.byref(nph), None) == 0
    assert nph.value == 0
    ref0 = o.WideFieldModelOracle((N, N, Nz), 0, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    ref0.Z = ref.Z                              # same basis as the handle still holds
    ref0.setModulus(BETA4)
    assert o.rel_l2(m.getPsf(), ref0.getPsf()) <= 1e-12
    assert o.rel_l2(m.getPsf(), psf_with) > 1e-3
    with pytest.raises(ValueError):
        m.apply_J_phase(o.synthetic_q(N, N, Nz))
    m.close()


def test_apply_j_all_checks_the_size_of_q(lib):
    ref, m = make_pair(32, 4, lib)
    with pytest.raises(ValueError, match="shape of the PSF"):
        m.apply_J_all(np.zeros((3, 32, 32)))
    m.close()


def test_eval_fg_keeps_the_mirror_in_step(lib):
    """After evalFG the mirror's parameterCoefs / ni / deltaX / deltaY are what setParam(x) leaves (WFM:412-422,
    1516-1531), so a later setNi / computeDefocus does not write stale values back."""
    N, Nz = 32, 32
    ref, m = make_pair(N, Nz, lib)
    f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, Nz), lib=lib)
    rng = np.random.default_rng(3)
    f.setPSF(rng.random((Nz, N, N)))
    f.setData(rng.random((Nz, N, N)))
    x = np.array([P["ni"] / P["lam"] * 1.001, 1.5e4, -0.5e4])
    cost, g = f.evalFG(m, m.DEFOCUS, x)
    assert np.isfinite(cost) and g.shape == (3,)
    np.testing.assert_array_equal(m.getDefocus(), x)
    assert m.getNi() == x[0] * P["lam"]
    np.testing.assert_array_equal(m.getPupilShift(), x[1:])
    np.testing.assert_array_equal(m.parameterCoefs[m.DEFOCUS].data, x)
    ref.setDefocus(x)
    m.computeDefocus()                           # writes lambda_ni / deltaX / deltaY of the MIRROR back to the device
    np.testing.assert_array_equal(m.getPsi(), ref.psi.ravel())
    a = o.synthetic_alpha(10) * 0.7
    f.evalFG(m, m.PHASE, a)
    np.testing.assert_array_equal(m.getPhaseCoefs().data, a)
    b = np.array([1.0, 0.2, -0.1, 0.05])
    f.evalFG(m, m.MODULUS, b)
    np.testing.assert_array_equal(m.getModulusCoefs().data, b)
    # and the one-call chain equals the explicit sequence setPhase -> getPsf -> cost/gradient -> apply_J_phase
    cost, g = f.evalFG(m, m.PHASE, a)
    m.setPhase(a)
    gq = np.zeros(N * N * Nz)
    c2 = f.computeCostAndGradient(1.0, m.getPsf(), gq, True)
    g2 = m.apply_J_phase(gq).data
    assert abs(cost - c2) <= 1e-13 * abs(c2) and o.rel_l2(g, g2) <= 1e-13
    f.close(); m.close()


def test_stream_ordering_entry_points(lib):
    ref, m = make_pair(32, 2, lib)
    m.waitForStream(0)
    m.orderStreamAfter(0)
    assert lib.wfm_wait_stream(None, None) != 0 and lib.wfm_fence_stream(None, None) != 0
    m.close()


def test_steady_state_set_phase_keeps_phi_and_strip_in_step(lib):
    """setPhase() on a packed strip takes the support-cell kernel (phi and its strip copy in one pass, no
    k_pack_strip): getPhi(), the PSF and the Jacobians must follow every new vector; a phase array loaded through the
    escape hatch (values off the support too) must be wiped by the next setPhase() exactly as the full pass does."""
    N, Nz = 32, 4
    ref, m = make_pair(N, Nz, lib)
    q = o.synthetic_q(N, N, Nz)
    m.getPsf()                                                   # packs the strip
    rng = np.random.default_rng(5)
    for it in range(3):
        a = rng.normal(0, 0.3, 10)
        n0 = lib.wfm_launch_count()
        m.setPhase(m.parameterSpace[m.PHASE].wrap(a.copy())); ref.setPhase(a)   # the optimiser's call: a vector of the space
        assert lib.wfm_launch_count() == n0 + 1                  # one setter launch ...
        m.computePsf()
        assert lib.wfm_launch_count() == n0 + 2                  # ... and the PSF pipeline: no k_pack_strip
        np.testing.assert_array_equal(m.getPhi(), ref.phi.ravel())
        assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
        assert o.rel_l2(m.apply_J_phase(q).data, ref.apply_J_phase(q)) <= 1e-12
    junk = rng.normal(size=N * N)                                # non-zero everywhere, also off the support
    assert lib.wfm_set_pupil_arrays(m.handle, None, junk.ctypes.data_as(C.c_void_p), None, None) == 0
    m.getPsf()                                                   # packs the strip again
    a = rng.normal(0, 0.3, 10)
    m.setPhase(a); ref.setPhase(a)
    np.testing.assert_array_equal(m.getPhi(), ref.phi.ravel())   # zero off the mask again
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    m.setModulus(BETA4); ref.setModulus(BETA4)                   # another setter: the strip is stale, full path
    m.setPhase(a * 0.5); ref.setPhase(a * 0.5)
    np.testing.assert_array_equal(m.getPhi(), ref.phi.ravel())
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    m.close()
