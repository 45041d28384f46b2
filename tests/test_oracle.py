"""Known-answer tests that pin the oracle.  The reference has no
tests of its own, so these follow from its formulas (WFM = WideFieldModel.java)."""
import numpy as np
import pytest

from oracle import wfm_oracle as o

P = o.DEFAULTS


def make(N=32, Nz=8, nPhase=10, nModulus=1, **kw):
    m = o.WideFieldModelOracle((N, N, Nz), nPhase, nModulus, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], **kw)
    if nPhase:
        m.setPhase(o.synthetic_alpha(nPhase))
    return m


def test_z_index_rule():
    # KAT 6: plane Nz/2 is positive (strict '>', WFM:302-309)
    s = o.defoc_scale(np.arange(32), 32, 1.0) / o.DEUXPI
    assert list(np.rint(s).astype(int)) == list(range(17)) + list(range(-15, 0))


def test_noll_indexing():
    # Noll (1976) table: j -> (n, m)
    want = {1: (0, 0), 2: (1, 1), 3: (1, 1), 4: (2, 0), 5: (2, 2), 6: (2, 2), 7: (3, 1), 8: (3, 1),
            9: (3, 3), 10: (3, 3), 11: (4, 0), 12: (4, 2), 13: (4, 2), 14: (4, 4), 15: (4, 4)}
    for j, nm in want.items():
        assert o.zernumero_noll(j) == nm


def test_radial_coefficients():
    # R_4^0 = 6r^4 - 6r^2 + 1 ; R_3^1 = 3r^3 - 2r
    np.testing.assert_allclose(o.coeff_radial(4, 0), [6, -6, 1], rtol=1e-13)
    np.testing.assert_allclose(o.coeff_radial(3, 1), [3, -2], rtol=1e-13)


def test_basis_orthonormal_and_supported_on_pupil():
    m = make(64, 4)
    G = m.Z @ m.Z.T
    assert np.abs(G - np.eye(m.Nzern)).max() < 1e-12
    assert not np.any(m.Z[:, ~m.mapPupil.ravel()])


def test_energy_parseval():
    # KAT 1: sum psf = sum rho^2 = 1 (WFM:284,327 with orthonormal Z, normalised beta)
    m = make(64, 32)
    psf = m.getPsf()
    assert abs((m.rho ** 2).sum() - 1) < 1e-13
    assert abs(psf.sum() - 1) < 1e-12
    np.testing.assert_allclose(psf.sum(axis=(1, 2)), 1.0 / 32, rtol=1e-12)


def test_focal_plane_and_conjugate_storage():
    # KAT 2 + 4
    m = make(32, 8)
    cpx = m.get_cpxPsf()
    a0 = np.fft.fft2(m.rho * np.exp(1j * m.phi))
    np.testing.assert_allclose(cpx[0, ..., 0], a0.real, atol=1e-15)
    np.testing.assert_allclose(cpx[0, ..., 1], -a0.imag, atol=1e-15)


def test_symmetry_alpha_zero():
    # KAT 3: alpha = 0 -> psf(iz) = psf(Nz-iz), psf(-x,-y) = psf(x,y)
    m = make(32, 8, nPhase=0)
    psf = m.getPsf()
    for iz in range(1, 4):
        np.testing.assert_allclose(psf[iz], psf[8 - iz], atol=1e-18)
    flipped = np.roll(psf[:, ::-1, ::-1], (1, 1), axis=(1, 2))
    np.testing.assert_allclose(psf, flipped, atol=1e-18)


def test_fft_against_longdouble_dft():
    m = make(32, 4)
    s = float(o.defoc_scale(3, 4, P["dz"]))
    A = m.rho * np.exp(1j * (m.phi + s * m.psi))
    ref = o.dft2_longdouble(A)
    cpx = m.get_cpxPsf()[3]
    got = cpx[..., 0] - 1j * cpx[..., 1]
    err = np.linalg.norm((got - ref).astype(np.complex128)) / np.linalg.norm(ref.astype(np.complex128))
    assert err < 5e-15


def _cost(m, q):
    return float(np.sum(q * m.getPsf()))


def test_phase_gradient_matches_finite_differences():
    # KAT 5: apply_J_phase is the exact gradient of C(alpha) = sum q*psf
    m = make(32, 8)
    q = o.synthetic_q(32, 32, 8)
    g = m.apply_J_phase(q)
    a0 = m.alpha.copy()
    for k in (0, 4, 9):
        h = 1e-6
        ap, am = a0.copy(), a0.copy()
        ap[k] += h
        am[k] -= h
        m.setPhase(ap); cp = _cost(m, q)
        m.setPhase(am); cm = _cost(m, q)
        fd = (cp - cm) / (2 * h)
        assert abs(fd - g[k]) <= 2e-6 * np.abs(g).max()
    m.setPhase(a0)


def test_defocus_gradient_is_half_finite_difference():
    # KAT 5 / Q3: live apply_J_defocus = 1/2 of the true gradient
    m = make(32, 8)
    q = o.synthetic_q(32, 32, 8)
    d = m.apply_J_defocus(q)
    base = [m.lambda_ni, m.deltaX, m.deltaY]
    for k, h in ((0, 1.0), (1, 1.0), (2, 1.0)):
        p, n = list(base), list(base)
        p[k] += h
        n[k] -= h
        m.setDefocus(p); cp = _cost(m, q)
        m.setDefocus(n); cm = _cost(m, q)
        fd = (cp - cm) / (2 * h)
        assert abs(0.5 * fd - d[k]) <= 1e-5 * np.abs(d).max() + 1e-7 * abs(d[k])
    m.setDefocus(base)


def test_modulus_pixel_gradient_and_quirk_modes():
    m = make(32, 8, nModulus=4)
    m.setModulus([1.0, 0.1, -0.05, 0.02])
    q = o.synthetic_q(32, 32, 8)
    gi = m.apply_J_modulus(q)
    m.modulus_mode = o.MODULUS_REFERENCE_LAST_PLANE
    gl = m.apply_J_modulus(q)
    assert gi.shape == gl.shape == (4,)
    assert not np.allclose(gi, gl)
    # last-plane mode == intended mode applied to a q that is zero except at iz = Nz-1
    q2 = np.zeros_like(q)
    q2[-1] = q[-1]
    m.modulus_mode = o.MODULUS_INTENDED
    np.testing.assert_allclose(m.apply_J_modulus(q2), gl, rtol=1e-12, atol=1e-20)


def test_linearity_and_shard_invariance():
    # KAT 7 + 8
    m = make(32, 8)
    q1 = o.synthetic_q(32, 32, 8, seed=1)
    q2 = o.synthetic_q(32, 32, 8, seed=2)
    g = m.apply_J_phase(2.0 * q1 - 3.0 * q2)
    np.testing.assert_allclose(g, 2 * m.apply_J_phase(q1) - 3 * m.apply_J_phase(q2), rtol=1e-10, atol=1e-18)
    cpx = m.get_cpxPsf()
    parts = []
    for z0 in (0, 4):
        c, p = o.compute_psf(m.rho, m.phi, m.psi, 8, P["dz"], z0=z0, nz_local=4)
        np.testing.assert_array_equal(c, cpx[z0:z0 + 4])
        parts.append(o.apply_J_phase(q1[z0:z0 + 4], c, m.rho, m.phi, m.psi, m.maskPupil, m.Z, 10, 8, P["dz"], z0=z0))
    np.testing.assert_allclose(parts[0] + parts[1], m.apply_J_phase(q1), rtol=1e-11, atol=1e-20)


def test_single_precision_mode_close_to_double():
    md = make(32, 8)
    ms = make(32, 8, single=True)
    assert ms.getPsf().dtype == np.float32
    assert o.rel_l2(ms.getPsf(), md.getPsf()) < 1e-5
    q = o.synthetic_q(32, 32, 8)
    assert o.rel_l2(ms.apply_J_phase(q.astype(np.float32)), md.apply_J_phase(q)) < 1e-4


def test_errors_mirror_reference():
    with pytest.raises(ValueError):
        o.WideFieldModelOracle((32, 16, 4), 10, 1, **P)          # WFM:158
    m = make(32, 4)
    with pytest.raises(ValueError):
        m.setDefocus([1.0, 2.0])                                  # Q4
    with pytest.raises(ValueError):
        m.setPhase(np.zeros(3))                                   # WFM:1629
    with pytest.raises(ValueError):
        m.apply_Jacobian(None, 7)                                 # WFM:407


def test_splitmix_counter_based():
    a = o.splitmix64_uniform(42, 0, 100)
    b = o.splitmix64_uniform(42, 50, 50)
    np.testing.assert_array_equal(a[50:], b)
    assert a.min() >= -1 and a.max() < 1
    # SplitMix64 known answer: first output for seed 0 is 0xE220A8397B1DCDAF
    z = np.uint64(0)
    u = o.splitmix64_uniform(0, 0, 1)[0]
    assert u == 2.0 * ((0xE220A8397B1DCDAF >> 11) / 2 ** 53) - 1.0


def test_blind_deconvolution_inner_loop_chain_rule():
    # config 3: d/d alpha of the data term 1/2||h(alpha)(*)obj - data||^2 == apply_J_phase(dcost/dh)
    m = make(32, 8)
    truth = make(32, 8)
    truth.setPhase(o.synthetic_alpha(10, seed=4321))
    O, data = o.bead_problem((8, 32, 32), truth.getPsf())
    a0 = m.alpha.copy()
    _, q = o.bead_cost_and_q(m.getPsf(), O, data)
    g = m.apply_J_phase(q)
    for k in (1, 6):
        h = 1e-6
        ap, am = a0.copy(), a0.copy()
        ap[k] += h
        am[k] -= h
        m.setPhase(ap); cp, _ = o.bead_cost_and_q(m.getPsf(), O, data)
        m.setPhase(am); cm, _ = o.bead_cost_and_q(m.getPsf(), O, data)
        assert abs((cp - cm) / (2 * h) - g[k]) <= 5e-6 * np.abs(g).max()


def test_weighted_convolution_cost_gradient_is_exact():
    # row f1: cost(h) = alpha/2 sum w (obj (*) h - y)^2 ; grad checked by central differences (the cost is
    # quadratic in h, so central differences are exact up to rounding) and against the unweighted bead helper
    rng = np.random.default_rng(5)
    shp = (4, 8, 8)
    obj, h, y = rng.normal(size=shp), rng.normal(size=shp), rng.normal(size=shp)
    w = rng.uniform(0.0, 2.0, size=shp)
    c0, g = o.weighted_convolution_cost(h, obj, y, w, alpha=0.7)
    for idx in [(0, 0, 0), (3, 7, 1), (2, 4, 5)]:
        e = np.zeros(shp); e[idx] = 1e-3
        cp, _ = o.weighted_convolution_cost(h + e, obj, y, w, alpha=0.7)
        cm, _ = o.weighted_convolution_cost(h - e, obj, y, w, alpha=0.7)
        assert abs((cp - cm) / 2e-3 - g[idx]) <= 1e-9 * np.abs(g).max()
    c1, g1 = o.weighted_convolution_cost(h, obj, y)
    c2, g2 = o.bead_cost_and_q(h, sfft_fftn(obj), y)
    assert abs(c1 - c2) <= 1e-12 * abs(c2) and o.rel_l2(g1, g2) <= 1e-13
    # direct (non-FFT) periodic convolution at one voxel
    k = (1, 2, 3)
    direct = sum(obj[a, b, c] * h[(k[0] - a) % 4, (k[1] - b) % 8, (k[2] - c) % 8]
                 for a in range(4) for b in range(8) for c in range(8))
    conv = np.fft.ifftn(np.fft.fftn(obj) * np.fft.fftn(h)).real
    assert abs(direct - conv[k]) <= 1e-12 * abs(direct)


def sfft_fftn(a):
    import scipy.fft
    return scipy.fft.fftn(a)
