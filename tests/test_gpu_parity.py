"""Parity tests proper: the sm_100a library, called through the C ABI, against the oracle, the
committed golden vectors, and size-independent properties at BASELINE's full sizes.
Tolerances are north_star's: rel-L2 <= 1e-12 (fp64), <= 1e-5 (fp32)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from oracle import wfm_oracle as o
from microtipi_b200 import _capi as capi, WideFieldModel
from tests.util import BETA4, P, gpu_lib, make_pair, oracle_basis, tol

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def lib():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    import __graft_entry__ as g
    g.build_library()
    return gpu_lib()


def _check_all(ref, m, N, Nz, single):
    t = tol(single)
    np.testing.assert_array_equal(m.getRho(), ref.rho.ravel())
    np.testing.assert_array_equal(m.getPhi(), ref.phi.ravel())
    np.testing.assert_array_equal(m.getPsi(), ref.psi.ravel())
    np.testing.assert_array_equal(m.getMaskPupil(), ref.maskPupil.ravel())
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= t
    assert o.rel_l2(m.get_cpxPsf(), ref.get_cpxPsf()) <= t
    q = o.synthetic_q(N, N, Nz, single=single)
    tj = 20 * t if single else t
    gp, gd, gm = ref.apply_J_phase(q), ref.apply_J_defocus(q), ref.apply_J_modulus(q)
    assert o.rel_l2(m.apply_J_phase(q).data, gp) <= tj
    assert o.rel_l2(m.apply_J_defocus(q).data, gd) <= tj
    assert o.rel_l2(m.apply_J_modulus(q).data, gm) <= tj
    d, p, mo = m.apply_J_all(q)
    assert o.rel_l2(np.concatenate([d, p, mo]), np.concatenate([gd, gp, gm])) <= tj
    m.setModulusMode(True)
    ref.modulus_mode = o.MODULUS_REFERENCE_LAST_PLANE
    assert o.rel_l2(m.apply_J_modulus(q).data, ref.apply_J_modulus(q)) <= tj


@pytest.mark.parametrize("N,Nz,single", [
    (32, 8, False), (64, 32, False),            # BASELINE config 1: 64x64x32 fp64
    (128, 9, False), (256, 16, False), (512, 6, False), (1024, 3, False), (2048, 1, False),
    (32, 8, True), (64, 32, True), (256, 16, True), (512, 6, True), (1024, 3, True),
])
def test_psf_and_jacobians_match_oracle(lib, N, Nz, single):
    ref, m = make_pair(N, Nz, lib, single=single)
    _check_all(ref, m, N, Nz, single)
    m.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))))
def test_against_committed_golden_vectors(lib, path):
    g = np.load(path)
    N, Nz, single = int(g["N"]), int(g["Nz"]), bool(g["single"])
    m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single, lib=lib,
                       basis=oracle_basis(N))
    m.setPhase(g["alpha"])
    m.setModulus(g["beta"])
    if np.any(g["delta"]):
        m.setDefocus([P["ni"] / P["lam"], g["delta"][0], g["delta"][1]])
    t = tol(single)
    assert o.rel_l2(m.getPsf(), g["psf"]) <= t
    cpx = m.get_cpxPsf()
    if "cpx_planes" in g:
        cpx = cpx[g["cpx_planes"]]
    assert o.rel_l2(cpx, g["cpx"]) <= t
    q = o.synthetic_q(N, N, Nz, single=single)
    tj = 20 * t if single else t
    assert o.rel_l2(m.apply_J_phase(q).data, g["j_phase"]) <= tj
    assert o.rel_l2(m.apply_J_defocus(q).data, g["j_defocus"]) <= tj
    assert o.rel_l2(m.apply_J_modulus(q).data, g["j_modulus"]) <= tj
    m.setModulusMode(True)
    assert o.rel_l2(m.apply_J_modulus(q).data, g["j_modulus_last_plane"]) <= tj
    m.close()


def test_device_side_basis(lib):
    for N in (64, 256):
        ref, m = make_pair(N, 2, lib, device_basis=True)
        Z = m.getZernike()
        assert o.rel_l2(Z, ref.Z) <= 1e-12
        assert np.abs(Z @ Z.T - np.eye(len(Z))).max() < 1e-12
        assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12      # whole constructor path on the device
        m.close()


def test_off_axis_and_state_protocol(lib):
    ref, m = make_pair(128, 5, lib, delta=(2e4, -2e4))
    assert m.PState == 0
    m.computePsf()
    n0 = lib.wfm_launch_count()
    m.computePsf()                                               # WFM:207: valid -> no kernels
    assert lib.wfm_launch_count() == n0
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    m.setDefocus([P["ni"] / P["lam"] * 1.01, 1e4, 3e4])
    ref.setDefocus([P["ni"] / P["lam"] * 1.01, 1e4, 3e4])
    assert m.PState == 0
    q = o.synthetic_q(128, 128, 5)
    assert o.rel_l2(m.apply_J_defocus(q).data, ref.apply_J_defocus(q)) <= 1e-12   # Q5: recompute
    np.testing.assert_array_equal(m.getPsi(), ref.psi.ravel())
    np.testing.assert_array_equal(m.getMaskPupil(), ref.maskPupil.ravel())
    m.close()


def test_edge_cases(lib):
    # single plane; odd slab at the wrap point (Nz/2 is positive, Nz/2+1 negative); zero gradient
    ref, m = make_pair(64, 1, lib)
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    m.close()
    N, Nz = 64, 9
    ref, full = make_pair(N, Nz, lib)
    q = o.synthetic_q(N, N, Nz)
    want = full.apply_J_all(q)
    acc = [np.zeros_like(w) for w in want]
    psf = []
    for z0, nzl in ((0, 4), (4, 1), (5, 4)):
        _, part = make_pair(N, Nz, lib, z0=z0, nz_local=nzl)
        psf.append(part.getPsf())
        for a, g in zip(acc, part.apply_J_all(q[z0:z0 + nzl])):
            a += g
        part.close()
    np.testing.assert_array_equal(np.concatenate(psf), full.getPsf())   # shard invariance (KAT 8)
    for a, w in zip(acc, want):
        assert o.rel_l2(a, w) <= 1e-12
    z = full.apply_J_all(np.zeros_like(q))
    assert all(not np.any(v) for v in z)
    # no phase coefficients (nPhase = 0): PSF symmetric in z and in (x, y)  (KAT 3)
    ref0, m0 = make_pair(64, 8, lib, nPhase=0, nModulus=1)
    psf0 = m0.getPsf()
    for iz in range(1, 4):
        np.testing.assert_allclose(psf0[iz], psf0[8 - iz], rtol=1e-12, atol=1e-20)
    with pytest.raises(ValueError):
        m0.apply_J_phase(q)
    m0.close()
    full.close()


def test_errors(lib):
    with pytest.raises(ValueError, match="Nx should equal Ny"):
        WideFieldModel((64, 32, 4), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], lib=lib)
    with pytest.raises(ValueError):
        WideFieldModel((5000, 5000, 4), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], lib=lib)
    ref, m = make_pair(32, 2, lib)
    with pytest.raises(ValueError):
        m.setDefocus([1.0, 2.0])
    with pytest.raises(ValueError, match="does not belong to any space"):
        m.apply_Jacobian(np.zeros((2, 32, 32)), object())
    m.close()


def test_full_size_properties_512x512x256(lib):
    """BASELINE's headline shape: size-independent properties instead of a full oracle run."""
    import torch
    N, Nz = 512, 256
    m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, False, lib=lib)
    alpha = o.synthetic_alpha(10)
    m.setPhase(alpha)
    m.setModulus(BETA4)
    rho, phi, psi = m.getRho(), m.getPhi(), m.getPsi()
    psf = m.getPsf()
    # KAT 1: per plane sum psf = sum rho^2 / Nz (Parseval with PSFnorm)
    e = float(np.sum(rho * rho))
    assert abs(e - 1.0) < 1e-12
    np.testing.assert_allclose(psf.sum(axis=(1, 2), dtype=np.float64), e / Nz, rtol=1e-12)
    # planes against the oracle on identical pupils (focal plane, wrap planes, last)
    for iz in (0, 1, 128, 129, 255):
        c, p = o.compute_psf(rho.reshape(N, N), phi.reshape(N, N), psi.reshape(N, N), Nz, P["dz"], z0=iz, nz_local=1)
        assert o.rel_l2(psf[iz], p[0]) <= 1e-12
    # device-resident Jacobian: linearity in q and agreement of a q supported on 2 planes with the oracle
    dev = torch.device("cuda:0")
    vox = N * N * Nz
    q1 = torch.empty(vox, dtype=torch.float64, device=dev)
    q2 = torch.empty(vox, dtype=torch.float64, device=dev)
    m.fillUniform(q1.data_ptr(), 42, 0, vox)
    m.fillUniform(q2.data_ptr(), 43, 0, vox)
    m.synchronize()
    np.testing.assert_array_equal(q1[:1000].cpu().numpy(), o.splitmix64_uniform(42, 0, 1000))
    L = m.gradLength()
    g = torch.zeros((3, L), dtype=torch.float64, device=dev)
    comb = 2.0 * q1 - 3.0 * q2
    torch.cuda.synchronize()                     # torch's stream produced the inputs; the handle runs on its own stream
    for i, qq in enumerate((q1, q2, comb)):
        m.applyJacobianDevice(7, qq.data_ptr(), g[i].data_ptr())
    m.synchronize()
    gh = g.cpu().numpy()
    assert o.rel_l2(gh[2], 2.0 * gh[0] - 3.0 * gh[1]) <= 1e-11
    qs = torch.zeros(vox, dtype=torch.float64, device=dev)
    sel = (3, 200)
    for iz in sel:
        qs[iz * N * N:(iz + 1) * N * N] = q1[iz * N * N:(iz + 1) * N * N]
    torch.cuda.synchronize()
    m.applyJacobianDevice(7, qs.data_ptr(), g[0].data_ptr())
    m.synchronize()
    got = g[0].cpu().numpy()
    Z = m.getZernike()
    mask = m.getMaskPupil().reshape(N, N)
    want_p = np.zeros(10)
    want_d = np.zeros(3)
    want_m = np.zeros(4)
    for iz in sel:
        qpl = q1[iz * N * N:(iz + 1) * N * N].cpu().numpy().reshape(1, N, N)
        c, _ = o.compute_psf(rho.reshape(N, N), phi.reshape(N, N), psi.reshape(N, N), Nz, P["dz"], z0=iz, nz_local=1)
        args = (qpl, c, rho.reshape(N, N), phi.reshape(N, N), psi.reshape(N, N), mask)
        want_p += o.apply_J_phase(*args, Z, 10, Nz, P["dz"], z0=iz)
        want_d += o.apply_J_defocus(*args, Nz, P["dz"], P["dxy"], P["ni"] / P["lam"], 0.0, 0.0, z0=iz)
        want_m += o.apply_J_modulus(*args, Z, BETA4, Nz, P["dz"], z0=iz)
    assert o.rel_l2(got[:3], want_d) <= 1e-12
    assert o.rel_l2(got[3:13], want_p) <= 1e-12
    assert o.rel_l2(got[13:], want_m) <= 1e-12
    m.close()


def test_config3_bead_volume_gradient(lib):
    """BASELINE config 3 at reduced size: q comes from an FFT-convolution data term on a synthetic bead
    (structured, strongly correlated with the PSF) instead of white noise."""
    N, Nz = 128, 16
    ref, m = make_pair(N, Nz, lib)
    truth = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    truth.setPhase(o.synthetic_alpha(10, seed=4321))
    truth.setModulus(BETA4)
    q = o.bead_gradient_q(ref.getPsf(), truth.getPsf())
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    for a, b in zip(m.apply_J_all(q), (ref.apply_J_defocus(q), ref.apply_J_phase(q), ref.apply_J_modulus(q))):
        assert o.rel_l2(a, b) <= 1e-12
    m.close()


@pytest.mark.parametrize("single", [False, True])
def test_config5_independent_models(lib, single):
    """BASELINE config 5 (batched PSF estimation over independent 256x256 bead PSFs) as a parity case:
    several models with parameter seeds 1234+b live side by side (distinct handles are independent)."""
    N, Nz, B = 256, 8, 3
    t = tol(single)
    models, refs = [], []
    for b in range(B):
        ref = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], single=single)
        m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single, lib=lib,
                           basis=oracle_basis(N))
        alpha = o.synthetic_alpha(10, seed=1234 + b)
        for mm in (ref, m):
            mm.setPhase(alpha)
            mm.setModulus(BETA4)
        models.append(m); refs.append(ref)
    for m in models:                                   # all PSFs first, then all Jacobians: handles interleave
        m.computePsf()
    for b, (m, ref) in enumerate(zip(models, refs)):
        q = o.synthetic_q(N, N, Nz, seed=42 + b, single=single)
        assert o.rel_l2(m.getPsf(), ref.getPsf()) <= t
        assert o.rel_l2(m.apply_J_phase(q).data, ref.apply_J_phase(q)) <= (20 * t if single else t)
    for m in models:
        m.close()


@pytest.mark.parametrize("N,Nz,B,single", [(64, 17, 3, False), (256, 8, 4, False), (256, 8, 3, True), (512, 3, 2, False)])
def test_config5_batch_handle(lib, N, Nz, B, single):
    """BASELINE config 5 on ONE batch handle (wfm_create_batch): every model has its own phase, modulus and
    defocus vector; PSF, cpxPsf and the three Jacobians of each model against its own oracle model."""
    from tests.test_emu_parity import _batch_case
    m, refs = _batch_case(lib, N, Nz, B, single)
    t = tol(single)
    tj = 20 * t if single else t
    rho, phi, psi, mask = m.getRho(), m.getPhi(), m.getPsi(), m.getMaskPupil()
    psf, cpx = m.getPsf(), m.get_cpxPsf()
    q = np.stack([o.synthetic_q(N, N, Nz, seed=42 + b, single=single) for b in range(B)])
    d, p, mo = m.applyJacobianBatch(q)
    for b, r in enumerate(refs):
        np.testing.assert_array_equal(rho[b], r.rho.ravel())
        np.testing.assert_array_equal(phi[b], r.phi.ravel())
        np.testing.assert_array_equal(psi[b], r.psi.ravel())
        np.testing.assert_array_equal(mask[b], r.maskPupil.ravel())
        assert o.rel_l2(psf[b], r.getPsf()) <= t
        assert o.rel_l2(cpx[b], r.get_cpxPsf()) <= t
        assert o.rel_l2(d[b], r.apply_J_defocus(q[b])) <= tj
        assert o.rel_l2(p[b], r.apply_J_phase(q[b])) <= tj
        assert o.rel_l2(mo[b], r.apply_J_modulus(q[b])) <= tj
    m.close()


def _bead_object(N, Nz, radius_px=2.5):
    zz, yy, xx = np.meshgrid(*(np.arange(n) - n // 2 for n in (Nz, N, N)), indexing="ij")
    obj = ((zz ** 2 + yy ** 2 + xx ** 2) <= radius_px ** 2).astype(np.float64)
    return np.roll(obj, (-(Nz // 2), -(N // 2), -(N // 2)), axis=(0, 1, 2))


@pytest.mark.parametrize("N,Nz", [(32, 32), (64, 128), (128, 64), (256, 32), (512, 32), (1024, 32)])
def test_convolution_data_term_matches_oracle(lib, N, Nz):
    """Row f1: TiPi WeightedConvolutionCost as PSF_Estimation.java:147-157,206 drives it (restated, unpinned)."""
    from microtipi_b200 import WeightedConvolutionCost, DoubleShapedVectorSpace
    rng = np.random.default_rng(11)
    shp = (Nz, N, N)
    obj, h, y = rng.normal(size=shp), rng.normal(size=shp), rng.normal(size=shp)
    w = rng.uniform(0.0, 2.0, size=shp)
    f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, Nz), lib=lib)
    f.setPSF(obj, (0, 0, 0)); f.setData(y); f.setWeights(w, True)
    g = np.zeros(h.size)
    c = f.computeCostAndGradient(0.7, h, g, True)
    c_ref, g_ref = o.weighted_convolution_cost(h, obj, y, w, 0.7)
    assert abs(c - c_ref) <= 1e-12 * abs(c_ref)
    assert o.rel_l2(g, g_ref) <= 1e-12
    c2 = f.computeCostAndGradient(0.7, h, g, False)
    assert c2 == c and o.rel_l2(g, 2 * g_ref) <= 1e-12
    f.setWeights(None)
    c3 = f.computeCostAndGradient(1.0, h, g, True)
    c3_ref, g3_ref = o.weighted_convolution_cost(h, obj, y)
    assert abs(c3 - c3_ref) <= 1e-12 * abs(c3_ref) and o.rel_l2(g, g3_ref) <= 1e-12
    f.close()


def test_eval_fg_inner_loop_matches_oracle_chain(lib):
    """BASELINE config 3 (reduced): one COMPUTE_FG evaluation of PSF_Estimation.fitPSF entirely on the device."""
    from microtipi_b200 import WeightedConvolutionCost, DoubleShapedVectorSpace
    N, Nz = 128, 32
    ref, m = make_pair(N, Nz, lib)
    truth = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    truth.setModulus(BETA4); truth.setPhase(o.synthetic_alpha(10, seed=4321))
    obj = _bead_object(N, Nz)
    _, data = o.bead_problem((Nz, N, N), truth.getPsf())
    f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, Nz), lib=lib)
    f.setPSF(obj); f.setData(data)
    for flag, x, refset, refj in (
            (m.PHASE, o.synthetic_alpha(10) + 0.01, ref.setPhase, ref.apply_J_phase),
            (m.DEFOCUS, np.array([P["ni"] / P["lam"], 1e4, -1e4]), ref.setDefocus, ref.apply_J_defocus),
            (m.MODULUS, np.array([1.0, 0.12, -0.04, 0.03]), ref.setModulus, ref.apply_J_modulus)):
        cost, g = f.evalFG(m, flag, x)
        refset(x)
        c_ref, q_ref = o.weighted_convolution_cost(ref.getPsf(), obj, data)
        assert abs(cost - c_ref) <= 1e-11 * abs(c_ref)
        assert o.rel_l2(g, refj(q_ref)) <= 1e-10
    f.close(); m.close()


def test_async_psf_readback_overlaps_and_orders(lib):
    """wfm_get_psf_async + wfm_wait_transfers: same bytes as getPsf(); the next computePsf is ordered after the copy."""
    N, Nz = 256, 32
    ref, m = make_pair(N, Nz, lib)
    nbytes = N * N * Nz * 8
    hp = C.c_void_p()
    assert lib.wfm_host_alloc(C.byref(hp), nbytes) == 0
    out = np.frombuffer((C.c_char * nbytes).from_address(hp.value), dtype=np.float64)
    q = o.synthetic_q(N, N, Nz)
    m.getPsfAsync(hp.value)
    g = m.apply_J_phase(q).data                 # H2D of q + Jacobian while the PSF is read back
    a2 = o.synthetic_alpha(10) * 0.5
    m.setPhase(a2); m.computePsf()              # must not overwrite the slab before the copy has finished
    m.waitTransfers()
    assert o.rel_l2(out, ref.getPsf()) <= 1e-12
    assert o.rel_l2(g, ref.apply_J_phase(q)) <= 1e-12
    ref.setPhase(a2)
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    lib.wfm_host_free(hp)
    m.close()


def chunked_host_case(lib, N, Nz, z0, nzl, single, env):
    """Host-buffer entry points with the slab moved in plane chunks (plane windows of the pipelines, copies on their own
    streams): getPsfAsync of a dirty PSF, then every host-q Jacobian, against the oracle; then the same through a clean
    PSF (single-piece read-back) -- the results must not depend on how the slab was cut."""
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ref = o.WideFieldModelOracle((N, N, 2), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], single=single)
        m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single, lib=lib,
                           basis=lambda nz: ref.Z[:nz], z0=z0, nz_local=nzl)
        alpha = o.synthetic_alpha(10)
        for mm in (ref, m):
            mm.setPhase(alpha)
            mm.setModulus(BETA4)
        q = o.synthetic_q(N, N, Nz, z0=z0, nz_local=nzl, single=single)
        cpx, psf, gd, gp, gm = _oracle_stack(ref, Nz, z0, nzl, q, single)
        t = tol(single)
        tj = 20 * t if single else t
        dt = np.float32 if single else np.float64
        nbytes = N * N * nzl * np.dtype(dt).itemsize
        hp = C.c_void_p()
        assert lib.wfm_host_alloc(C.byref(hp), nbytes) == 0
        out = np.frombuffer((C.c_char * nbytes).from_address(hp.value), dtype=dt).reshape(nzl, N, N)
        out[:] = -1
        m.getPsfAsync(hp.value)                       # dirty PSF: windows of computePsf, each read back behind it
        g = m.apply_J_phase(q).data                   # q in chunks on its own stream, adjoint windows behind them
        m.waitTransfers()
        assert max(o.rel_l2(out[l], psf[l]) for l in range(nzl)) <= t
        assert o.rel_l2(m.get_cpxPsf(), cpx) <= t
        assert o.rel_l2(g, gp) <= tj
        d, p, mo = m.apply_J_all(q)
        assert o.rel_l2(d, gd) <= tj and o.rel_l2(p, gp) <= tj and o.rel_l2(mo, gm) <= tj
        assert o.rel_l2(m.apply_J_defocus(q).data, gd) <= tj
        assert o.rel_l2(m.apply_J_modulus(q).data, gm) <= tj
        m.setPhase(alpha)                             # dirty again: the synchronous getPsf() (plane windows; staged through
        got = m.getPsf()                              # pinned slots by the host threads when the array is pageable)
        assert max(o.rel_l2(got[l], psf[l]) for l in range(nzl)) <= t
        out[:] = -1
        m.getPsfAsync(hp.value); m.waitTransfers()    # clean PSF: one copy
        assert max(o.rel_l2(out[l], psf[l]) for l in range(nzl)) <= t
        m.setPhase(alpha * 0.5)                       # dirty again; Jacobian first (quirk Q5 recomputes the whole slab)
        ref.setPhase(alpha * 0.5)
        cpx2, psf2, gd2, gp2, gm2 = _oracle_stack(ref, Nz, z0, nzl, q, single)
        assert o.rel_l2(m.apply_J_phase(q).data, gp2) <= tj
        assert o.rel_l2(m.getPsf(), psf2) <= t
        lib.wfm_host_free(hp)
        m.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("N,Nz,z0,nzl,single,env", [
    (512, 256, 0, 256, False, {}),                                                        # default policy: 4 chunks of 64 planes
    (256, 128, 40, 70, False, {"WFM_HOST_CHUNKS": "4", "WFM_HOST_CHUNK_MIN_BYTES": "1"}),  # ragged last chunk (18, 18, 18, 16)
    (256, 64, 0, 64, True, {"WFM_HOST_CHUNKS": "3", "WFM_HOST_CHUNK_MIN_BYTES": "1"}),
    (512, 256, 0, 256, False, {"WFM_FORCE_STAGED": "1"}),                                 # pageable-array path, default policy
    (256, 128, 40, 70, False, {"WFM_FORCE_STAGED": "1", "WFM_STAGED_MIN_BYTES": "1", "WFM_HOST_THREADS": "3",
                               "WFM_HOST_SHARE_BYTES": "300000", "WFM_HOST_CHUNKS": "4", "WFM_HOST_CHUNK_MIN_BYTES": "1"}),
    (128, 16, 0, 16, True, {"WFM_FORCE_STAGED": "1", "WFM_STAGED_MIN_BYTES": "1", "WFM_HOST_THREADS": "1",
                            "WFM_HOST_SHARE_BYTES": "70000"}),                           # one window, one thread
])
def test_host_paths_in_plane_chunks(lib, N, Nz, z0, nzl, single, env):
    chunked_host_case(lib, N, Nz, z0, nzl, single, env)


@pytest.mark.parametrize("N,Nz", [(64, 32), (256, 64)])
def test_rolled_psf_and_mtf(lib, N, Nz):
    """Row f4: ArrayUtils.roll(getPsf()) (BlindDeconvJob.java:100) and the intended getMtf() (WFM:1807-1828)."""
    ref, m = make_pair(N, Nz, lib)
    psf = ref.getPsf()
    np.testing.assert_array_equal(m.getPsfRolled(), o.roll_psf(m.getPsf()))
    assert o.rel_l2(m.getPsfRolled(), o.roll_psf(psf)) <= 1e-12
    assert o.rel_l2(m.getMtf(), o.mtf(psf)) <= 1e-12
    m.close()


def test_generic_kernels_when_the_pupil_is_wide(lib):
    """A pupil wider than N/4 switches the pruned ("narrow") kernels off: both variants must agree with the oracle."""
    N, Nz = 128, 8
    big = dict(P); big["dxy"] = 2.2 * P["dxy"]                     # pupil radius 0.37 N > N/4
    ref = o.WideFieldModelOracle((N, N, Nz), 10, 4, big["NA"], big["lam"], big["ni"], big["dxy"], big["dz"])
    m = WideFieldModel((N, N, Nz), 10, 4, big["NA"], big["lam"], big["ni"], big["dxy"], big["dz"], False, False, lib=lib,
                       basis=lambda nz: o.compute_zernike(nz, N, N, big["NA"], big["lam"], big["dxy"]))
    for mm in (ref, m):
        mm.setPhase(o.synthetic_alpha(10)); mm.setModulus(BETA4)
    assert m.activeExtent()[0] > N // 2
    q = o.synthetic_q(N, N, Nz)
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    for a, b in zip(m.apply_J_all(q), (ref.apply_J_defocus(q), ref.apply_J_phase(q), ref.apply_J_modulus(q))):
        assert o.rel_l2(a, b) <= 1e-12
    m.close()


# ---- full-stack parity at the BASELINE shapes (every plane against the oracle) ---------------------------------
# These are the only cases where the plane index exceeds the intermediate ring of the pipelines (ring = 2*lag + 2
# planes, ~44 at 512^2), i.e. where ring slots are re-used and the B(p - ring) -> A(p) dependency is live.
def _oracle_stack(ref, Nz, z0, nzl, q, single=False):
    """Oracle PSF / cpxPsf / three Jacobians of the slab [z0, z0+nzl) of a global stack of Nz planes."""
    w = os.cpu_count() or 1
    rho, phi, psi, mask = ref.rho, ref.phi, ref.psi, ref.maskPupil
    cpx, psf = o.compute_psf(rho, phi, psi, Nz, P["dz"], single=single, z0=z0, nz_local=nzl, workers=w)
    gp = o.apply_J_phase(q, cpx, rho, phi, psi, mask, ref.Z, ref.nPhase, Nz, P["dz"], single=single, z0=z0, workers=w)
    gd = o.apply_J_defocus(q, cpx, rho, phi, psi, mask, Nz, P["dz"], P["dxy"], ref.lambda_ni, ref.deltaX, ref.deltaY,
                           single=single, z0=z0, workers=w)
    gm = o.apply_J_modulus(q, cpx, rho, phi, psi, mask, ref.Z, ref.beta, Nz, P["dz"], single=single, z0=z0, workers=w)
    return cpx, psf, gd, gp, gm


@pytest.mark.parametrize("N,Nz,z0,nzl,single", [
    (512, 256, 0, 256, False),        # BASELINE config 3 / the headline shape: all 256 planes
    (256, 128, 0, 128, False),        # BASELINE config 2
    (1024, 512, 448, 64, False),      # BASELINE config 4: the last GPU's slab of the 1024^2 x 512 stack
    (512, 256, 64, 96, True),         # optional fp32 mode, a slab that straddles the z wrap (Nz/2 = 128)
])
def test_full_stack_parity_at_baseline_shapes(lib, N, Nz, z0, nzl, single):
    full_stack_case(lib, N, Nz, z0, nzl, single)


def full_stack_case(lib, N, Nz, z0, nzl, single):
    ref = o.WideFieldModelOracle((N, N, 2), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], single=single)
    m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single, lib=lib,
                       basis=lambda nz: ref.Z[:nz], z0=z0, nz_local=nzl)
    alpha = o.synthetic_alpha(10)
    for mm in (ref, m):
        mm.setPhase(alpha)
        mm.setModulus(BETA4)
    q = o.synthetic_q(N, N, Nz, z0=z0, nz_local=nzl, single=single)
    cpx, psf, gd, gp, gm = _oracle_stack(ref, Nz, z0, nzl, q, single)
    t = tol(single)
    tj = 20 * t if single else t
    got_psf, got_cpx = m.getPsf(), m.get_cpxPsf()
    assert o.rel_l2(got_psf, psf) <= t
    assert o.rel_l2(got_cpx, cpx) <= t
    worst = max(o.rel_l2(got_psf[l], psf[l]) for l in range(nzl))          # no single plane may hide in the norm
    assert worst <= t, f"worst plane {worst:.2e}"
    worst = max(o.rel_l2(got_cpx[l], cpx[l]) for l in range(nzl))
    assert worst <= t, f"worst cpx plane {worst:.2e}"
    d, p, mo = m.apply_J_all(q)
    assert o.rel_l2(d, gd) <= tj and o.rel_l2(p, gp) <= tj and o.rel_l2(mo, gm) <= tj
    assert o.rel_l2(m.apply_J_phase(q).data, gp) <= tj                    # the single-Jacobian launch (kinds = 2)
    # a q that is non-zero on one late plane only: an error in that plane's ring slot cannot average out
    for l in sorted({nzl - 1, nzl // 2 + 1, min(nzl - 1, 47)}):
        q1 = np.zeros_like(q)
        q1[l] = q[l]
        want = o.apply_J_phase(q1[l:l + 1], cpx[l:l + 1], ref.rho, ref.phi, ref.psi, ref.maskPupil, ref.Z, 10, Nz,
                               P["dz"], single=single, z0=z0 + l)
        assert o.rel_l2(m.apply_J_phase(q1).data, want) <= tj, f"plane {l}"
    m.close()


@pytest.mark.parametrize("lag", ["1", "2", "3"])
def test_tight_ring_slot_reuse(lib, monkeypatch, lag):
    """A ring of 4 / 6 / 8 planes (WFM_PIPE_LAG): the producers of a plane wait on the counter of the slot's previous
    tenant for most items, so a slot is rewritten microseconds after its consumer dropped its lines from L2
    (wfm_discard_l2) and published it -- the ordering the discard relies on, under the least slack the queue allows."""
    monkeypatch.setenv("WFM_PIPE_LAG", lag)
    for _ in range(2):
        full_stack_case(lib, 512, 96, 16, 64, False)
    full_stack_case(lib, 256, 64, 0, 64, True)


@pytest.mark.parametrize("single", [False, True])
def test_config5_batch_of_eight_256x256x64(lib, single):
    """BASELINE config 5 at its own plane count: 8 of the 64 models (512 planes through one pipeline launch pair)."""
    from tests.test_emu_parity import _batch_case
    N, Nz, B = 256, 64, 8
    m, refs = _batch_case(lib, N, Nz, B, single)
    t = tol(single)
    tj = 20 * t if single else t
    psf, cpx = m.getPsf(), m.get_cpxPsf()
    q = np.stack([o.synthetic_q(N, N, Nz, seed=42 + b, single=single) for b in range(B)])
    d, p, mo = m.applyJacobianBatch(q)
    for b, r in enumerate(refs):
        assert o.rel_l2(psf[b], r.getPsf()) <= t
        assert o.rel_l2(cpx[b], r.get_cpxPsf()) <= t
        assert o.rel_l2(d[b], r.apply_J_defocus(q[b])) <= tj
        assert o.rel_l2(p[b], r.apply_J_phase(q[b])) <= tj
        assert o.rel_l2(mo[b], r.apply_J_modulus(q[b])) <= tj
    m.close()


def test_escape_hatch_identical_pupils_on_the_gpu(lib):
    """wfm_set_pupil_arrays (SURVEY 8 b3 "identical synthetic pupils"): arbitrary rho / phi / psi / mask -- not the
    output of any setter, support wider than the optical mask -- through the generic and the narrow kernels."""
    escape_hatch_case(lib, ((128, 40, 0.2), (256, 24, 0.4)))         # 0.2 N < N/4: narrow kernels; 0.4 N: generic


def escape_hatch_case(lib, cases):
    for N, Nz, rad in cases:
        rng = np.random.default_rng(N)
        ky, kx = np.meshgrid(o.kappa(N), o.kappa(N), indexing="ij")
        disk = (kx * kx + ky * ky) < (rad * N) ** 2
        rho = np.where(disk, rng.uniform(0.2, 1.0, (N, N)), 0.0)
        rho /= np.sqrt(np.sum(rho * rho))
        phi = np.where(disk, rng.normal(0.0, 1.0, (N, N)), 0.0)
        psi = np.where(disk, rng.uniform(1e6, 3e6, (N, N)), 0.0)
        mask = disk & (rng.random((N, N)) > 0.05)                     # a few support pixels are off the mask
        ref = o.WideFieldModelOracle((N, N, 2), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
        m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, False, lib=lib,
                           basis=lambda nz: ref.Z[:nz])
        m.setPupilArrays(rho, phi, psi, mask)
        ref.rho, ref.phi, ref.psi, ref.maskPupil = rho, phi, psi, mask
        ref.nPhase = 10
        ref.beta = np.array([1.0, 0.0, 0.0, 0.0])                     # what the constructor left on the handle
        q = o.synthetic_q(N, N, Nz)
        cpx, psf, gd, gp, gm = _oracle_stack(ref, Nz, 0, Nz, q)
        np.testing.assert_array_equal(m.getRho(), rho.ravel())
        np.testing.assert_array_equal(m.getMaskPupil(), mask.ravel())
        assert o.rel_l2(m.getPsf(), psf) <= 1e-12
        assert o.rel_l2(m.get_cpxPsf(), cpx) <= 1e-12
        d, p, mo = m.apply_J_all(q)
        assert o.rel_l2(d, gd) <= 1e-12 and o.rel_l2(p, gp) <= 1e-12 and o.rel_l2(mo, gm) <= 1e-12
        m.close()


def test_handle_on_a_fresh_thread_keeps_its_device(lib):
    """Every ABI entry point makes the handle's device current itself (a new host thread starts on device 0) and
    hands the caller's device back."""
    import threading
    import torch
    dev = torch.cuda.device_count() - 1                               # device 1 when the box has more than one GPU
    ref, m = make_pair(64, 8, lib, device=dev)
    q = o.synthetic_q(64, 64, 8)
    out = {}

    def worker():
        out["before"] = torch.cuda.current_device()
        out["psf"] = m.getPsf()
        out["g"] = m.apply_J_phase(q).data
        out["after"] = torch.cuda.current_device()
    torch.cuda.set_device(0)
    th = threading.Thread(target=worker)
    th.start(); th.join()
    assert out["before"] == out["after"] == 0
    assert o.rel_l2(out["psf"], ref.getPsf()) <= 1e-12
    assert o.rel_l2(out["g"], ref.apply_J_phase(q)) <= 1e-12
    m.close()


def test_multi_device_handle(lib):
    """wfm_create_multi on min(2, device_count) GPUs: the reference-facing calls (setters broadcast, getPsf gathered
    slab by slab, apply_J_* with q split across the devices) and the device-resident path whose partial K-vectors
    travel over peer memory to the first device."""
    import torch
    from tests.test_multi_device import multi_case
    n = min(2, torch.cuda.device_count())
    N, Nz = 256, 41
    ref, m, q = multi_case(lib, N, Nz, list(range(n)))
    m.setModulusMode(False)
    ref.modulus_mode = o.MODULUS_INTENDED
    slabs = [torch.from_numpy(np.ascontiguousarray(q[z0:z0 + nz])).to(torch.device("cuda", d)) for (d, z0, nz, _) in m.parts()]
    grad = torch.zeros(m.gradLength(), dtype=torch.float64, device="cuda:0")
    torch.cuda.synchronize()
    for _ in range(3):                                                 # repeated calls re-use the landing slots
        m.applyJacobianDeviceMulti(7, [s.data_ptr() for s in slabs], grad.data_ptr())
    m.synchronize()
    want = np.concatenate([ref.apply_J_defocus(q), ref.apply_J_phase(q), ref.apply_J_modulus(q)])
    assert o.rel_l2(grad.cpu().numpy(), want) <= 1e-12
    m.close()


def test_z_sharded_data_term_and_inner_loop(lib):
    """Row f1 across GPUs: wfm_conv_create_multi + wfm_eval_fg on min(2, device_count) devices (slabs <-> pencils
    through peer stores) against the oracle."""
    import torch
    from tests.test_multi_device import conv_multi_case
    n = min(2, torch.cuda.device_count())
    conv_multi_case(lib, 128, 64, list(range(n)))
    conv_multi_case(lib, 64, 128, list(range(n)))


@pytest.mark.parametrize("N,Nz,single", [(100, 9, False), (384, 5, False), (48, 6, True), (1000, 2, False)])
def test_any_n_path_matches_oracle(lib, N, Nz, single):
    """Sizes off the pipeline plans (the reference's JTransforms takes any Nx == Ny, WFM:319): wfm_generic.cuh."""
    ref, m = make_pair(N, Nz, lib, single=single)
    _check_all(ref, m, N, Nz, single)
    m.close()
