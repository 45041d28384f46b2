"""CPU-side checks of the kernel sources and of the C-ABI host logic.

The SAME .cu/.cuh files nvcc builds for sm_100a are compiled with g++ against the test-only
emulator tests/emu/cuda_emu.h (one fiber per CUDA thread) and compared with the oracle.  This
validates index math, shared-memory exchanges, barriers and the PState state machine in a
container without a GPU; the parity tests proper are tests/test_gpu_parity.py (-m gpu)."""
import ctypes as C

import numpy as np
import pytest

from oracle import wfm_oracle as o
from microtipi_b200 import _capi as capi, WideFieldModel
from tests.util import P, emu_lib, make_pair, tol


@pytest.fixture(scope="module")
def lib():
    return emu_lib()


@pytest.mark.parametrize("N,Nz,single", [(32, 6, False), (64, 4, False), (128, 3, False), (64, 4, True),
                                         (256, 2, False), (512, 2, False), (512, 2, True), (1024, 1, False),   # production-size plans (narrow, affine)
                                         (128, 3, True), (256, 2, True), (1024, 1, True)])                       # fp32 rows with per-block offsets
def test_psf_and_jacobians_match_oracle(lib, N, Nz, single):
    ref, m = make_pair(N, Nz, lib, single=single)
    t = tol(single)
    np.testing.assert_array_equal(m.getRho(), ref.rho.ravel())       # elementwise setters are bit exact
    np.testing.assert_array_equal(m.getPhi(), ref.phi.ravel())
    np.testing.assert_array_equal(m.getPsi(), ref.psi.ravel())
    np.testing.assert_array_equal(m.getMaskPupil(), ref.maskPupil.ravel())
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= t
    assert o.rel_l2(m.get_cpxPsf(), ref.get_cpxPsf()) <= t
    q = o.synthetic_q(N, N, Nz, single=single)
    tj = 20 * t if single else t
    assert o.rel_l2(m.apply_J_phase(q).data, ref.apply_J_phase(q)) <= tj
    assert o.rel_l2(m.apply_J_defocus(q).data, ref.apply_J_defocus(q)) <= tj
    assert o.rel_l2(m.apply_J_modulus(q).data, ref.apply_J_modulus(q)) <= tj
    d, p, mo = m.apply_J_all(q)
    assert o.rel_l2(np.concatenate([d, p, mo]),
                    np.concatenate([ref.apply_J_defocus(q), ref.apply_J_phase(q), ref.apply_J_modulus(q)])) <= tj
    m.setModulusMode(True)                                           # quirk Q1 compat mode
    ref.modulus_mode = o.MODULUS_REFERENCE_LAST_PLANE
    assert o.rel_l2(m.apply_J_modulus(q).data, ref.apply_J_modulus(q)) <= tj
    m.close()


def test_device_side_basis_matches_oracle(lib):
    ref, m = make_pair(32, 2, lib, device_basis=True)
    assert o.rel_l2(m.getZernike(), ref.Z) <= 1e-12
    G = m.getZernike() @ m.getZernike().T
    assert np.abs(G - np.eye(len(G))).max() < 1e-12
    m.close()


def test_off_axis_defocus_and_dirty_state_protocol(lib):
    ref, m = make_pair(32, 5, lib, delta=(2e4, -2e4))
    assert m.PState == 0
    m.computePsf()
    assert m.PState == 1 and lib.wfm_psf_state(m.handle) == 1
    n0 = lib.wfm_launch_count()
    m.computePsf()                                                   # valid -> no work (WFM:207)
    assert lib.wfm_launch_count() == n0
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    m.setPhase(np.zeros(10))                                         # every setter ends in freeMem()
    assert m.PState == 0 and lib.wfm_psf_state(m.handle) == 0
    ref.setPhase(np.zeros(10))
    q = o.synthetic_q(32, 32, 5)
    g = m.apply_J_defocus(q).data                                    # quirk Q5: dirty -> recompute first
    assert m.PState == 1
    assert o.rel_l2(g, ref.apply_J_defocus(q)) <= 1e-12
    m.close()


def test_z_slabs_sum_to_full_gradient(lib):
    N, Nz = 32, 7
    ref, full = make_pair(N, Nz, lib)
    q = o.synthetic_q(N, N, Nz)
    want = full.apply_J_all(q)
    acc = [np.zeros_like(w) for w in want]
    psf = []
    for z0, nzl in ((0, 3), (3, 4)):
        _, part = make_pair(N, Nz, lib, z0=z0, nz_local=nzl)
        psf.append(part.getPsf())
        for a, g in zip(acc, part.apply_J_all(q[z0:z0 + nzl])):
            a += g
        part.close()
    np.testing.assert_array_equal(np.concatenate(psf), full.getPsf())
    for a, w in zip(acc, want):
        assert o.rel_l2(a, w) <= 1e-12
    full.close()


def test_escape_hatch_identical_pupils(lib):
    N, Nz = 32, 4
    ref, _m = make_pair(N, Nz, lib)
    _m.close()
    m = WideFieldModel((N, N, Nz), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, False, lib=lib,
                       basis=lambda nz: ref.Z[:nz])
    m.setPupilArrays(ref.rho, ref.phi, ref.psi, ref.maskPupil)
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    q = o.synthetic_q(N, N, Nz)
    assert o.rel_l2(m.apply_J_phase(q).data, ref.apply_J_phase(q)) <= 1e-12
    m.close()


def test_error_behaviour_mirrors_reference(lib):
    with pytest.raises(ValueError, match="Nx should equal Ny"):       # WFM:158
        WideFieldModel((32, 64, 4), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], lib=lib)
    with pytest.raises(ValueError):                                   # beyond the any-N path (N <= 4096)
        WideFieldModel((5000, 5000, 4), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], lib=lib)
    ref, m = make_pair(32, 2, lib)
    with pytest.raises(ValueError):                                   # quirk Q4: length-2 defocus
        m.setDefocus([1.0, 2.0])
    two = (C.c_double * 2)(1.0, 2.0)
    assert lib.wfm_set_defocus(m.handle, two, 2) == capi.WFM_ERR_INVALID_ARG
    assert b"bad defocus" in lib.wfm_last_error(m.handle)             # WFM:1530
    with pytest.raises(ValueError):                                   # WFM:1629
        m.setPhase(m.parameterSpace[m.MODULUS].create(0.0))
    with pytest.raises(ValueError, match="does not belong to any space"):   # WFM:407
        m.apply_Jacobian(np.zeros((2, 32, 32)), object())
    g = m.apply_Jacobian(o.synthetic_q(32, 32, 2), m.parameterSpace[m.PHASE])
    assert g.getNumber() == 10 and g.belongsTo(m.parameterSpace[m.PHASE])
    h = m.handle
    out = (C.c_double * 10)()
    assert lib.wfm_apply_j_phase(h, None, out, 10) == capi.WFM_ERR_INVALID_ARG
    assert lib.wfm_apply_j_phase(h, out, out, 9) == capi.WFM_ERR_INVALID_ARG
    assert b"nPhase" in lib.wfm_last_error(h)
    m.close()


def test_fill_uniform_matches_oracle_generator(lib):
    ref, m = make_pair(32, 2, lib)
    n = 1000
    buf = np.zeros(n)
    m.fillUniform(buf.ctypes.data, seed=42, first_index=12345, count=n)   # emulated "device" memory is host memory
    np.testing.assert_array_equal(buf, o.splitmix64_uniform(42, 12345, n))
    m.close()


@pytest.mark.parametrize("N,Nz", [(32, 32), (64, 32), (32, 64), (128, 32), (64, 128), (256, 32), (512, 32)])
def test_convolution_data_term_matches_oracle(lib, N, Nz):
    # row f1 (TiPi WeightedConvolutionCost as PSF_Estimation.java:147-157 drives it): 3-D FFT convolution cost
    # and gradient through the same kernel sources, with weights, alpha and the clr flag
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
    c2 = f.computeCostAndGradient(0.7, h, g, False)                 # clr = false accumulates
    assert c2 == c and o.rel_l2(g, 2 * g_ref) <= 1e-12
    f.setWeights(None)
    c3 = f.computeCostAndGradient(1.0, h, g, True)
    c3_ref, g3_ref = o.weighted_convolution_cost(h, obj, y)
    assert abs(c3 - c3_ref) <= 1e-12 * abs(c3_ref) and o.rel_l2(g, g3_ref) <= 1e-12
    with pytest.raises(ValueError):
        f.setPSF(obj, (1, 0, 0))
    with pytest.raises(ValueError):
        WeightedConvolutionCost.build(DoubleShapedVectorSpace(32, 64, 32), lib=lib)
    f.close()


def test_eval_fg_one_call_inner_loop(lib):
    # wfm_eval_fg == setParam -> computePsf -> computeCostAndGradient -> apply_Jacobian (PSF_Estimation.java:202-217)
    from microtipi_b200 import WeightedConvolutionCost, DoubleShapedVectorSpace
    N, Nz = 32, 32
    ref, m = make_pair(N, Nz, lib)
    truth = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    truth.setModulus([1.0, 0.1, -0.05, 0.02]); truth.setPhase(o.synthetic_alpha(10, seed=4321))
    zz, yy, xx = np.meshgrid(*(np.arange(n) - n // 2 for n in (Nz, N, N)), indexing="ij")
    obj = np.roll(((zz ** 2 + yy ** 2 + xx ** 2) <= 2.5 ** 2).astype(np.float64), (-(Nz // 2), -(N // 2), -(N // 2)), (0, 1, 2))
    _, data = o.bead_problem((Nz, N, N), truth.getPsf())
    f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, Nz), lib=lib)
    f.setPSF(obj); f.setData(data)
    x = o.synthetic_alpha(10) + 0.01
    cost, g = f.evalFG(m, m.PHASE, x)
    ref.setPhase(x)
    c_ref, q_ref = o.weighted_convolution_cost(ref.getPsf(), obj, data)
    g_ref = ref.apply_J_phase(q_ref)
    assert abs(cost - c_ref) <= 1e-11 * abs(c_ref)
    assert o.rel_l2(g, g_ref) <= 1e-10          # conditioning of the chain (residual is a difference of near-equal volumes)
    d = np.array([P["ni"] / P["lam"], 1e4, -1e4])
    cost, g = f.evalFG(m, m.DEFOCUS, d)
    ref.setDefocus(d)
    c_ref, q_ref = o.weighted_convolution_cost(ref.getPsf(), obj, data)
    assert abs(cost - c_ref) <= 1e-11 * abs(c_ref)
    assert o.rel_l2(g, ref.apply_J_defocus(q_ref)) <= 1e-10
    f.close(); m.close()


def test_rolled_psf_and_mtf(lib):
    # row f4: ArrayUtils.roll(getPsf()) (BlindDeconvJob.java:100) and the intended getMtf() (WFM:1807-1828)
    ref, m = make_pair(32, 32, lib)
    psf = ref.getPsf()
    np.testing.assert_array_equal(m.getPsfRolled(), o.roll_psf(m.getPsf()))
    assert o.rel_l2(m.getPsfRolled(), o.roll_psf(psf)) <= 1e-12
    assert o.rel_l2(m.getMtf(), o.mtf(psf)) <= 1e-12
    assert abs(m.getMtf()[0, 0, 0, 0] - psf.sum()) <= 1e-12          # DC term = total energy (= sum rho^2)
    m.close()
    sl = WideFieldModel((32, 32, 32), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, False, lib=lib,
                        z0=8, nz_local=8)
    with pytest.raises(RuntimeError):
        sl.getPsfRolled()                                           # z roll crosses slabs
    sl.close()


def test_wide_pupil_uses_generic_kernels(lib):
    # pupil radius > N/4: the "narrow" (pruned first-stage) kernels must not be selected; both paths match the oracle
    N, Nz = 64, 3
    big = dict(P); big["dxy"] = 2.2 * P["dxy"]
    ref = o.WideFieldModelOracle((N, N, Nz), 10, 4, big["NA"], big["lam"], big["ni"], big["dxy"], big["dz"])
    m = WideFieldModel((N, N, Nz), 10, 4, big["NA"], big["lam"], big["ni"], big["dxy"], big["dz"], False, False, lib=lib,
                       basis=lambda nz: o.compute_zernike(nz, N, N, big["NA"], big["lam"], big["dxy"]))
    for mm in (ref, m):
        mm.setPhase(o.synthetic_alpha(10)); mm.setModulus([1.0, 0.1, -0.05, 0.02])
    assert m.activeExtent()[0] > N // 2
    q = o.synthetic_q(N, N, Nz)
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= 1e-12
    assert o.rel_l2(m.apply_J_phase(q).data, ref.apply_J_phase(q)) <= 1e-12
    assert o.rel_l2(m.apply_J_defocus(q).data, ref.apply_J_defocus(q)) <= 1e-12
    m.close()


@pytest.mark.parametrize("N,single", [(512, False), (512, True), (256, False)])
def test_generic_kernels_at_production_size(lib, monkeypatch, N, single):
    # WFM_NO_NARROW=1 forces the un-pruned pipelines (the ones a wide pupil selects) at the sizes whose row layout is
    # affine (512) and is not (256): same results as the narrow kernels and as the oracle
    monkeypatch.setenv("WFM_NO_NARROW", "1")
    Nz = 2
    ref, m = make_pair(N, Nz, lib, single=single)
    t = tol(single)
    q = o.synthetic_q(N, N, Nz, single=single)
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= t
    d, p, mo = m.apply_J_all(q)
    assert o.rel_l2(np.concatenate([d, p, mo]), np.concatenate([ref.apply_J_defocus(q), ref.apply_J_phase(q),
                                                                ref.apply_J_modulus(q)])) <= (20 * t if single else t)
    m.close()


def _batch_case(lib, N, Nz, B, single, nModulus=4):
    """B models with different phase / modulus / defocus vectors on one batch handle vs B oracle models."""
    from microtipi_b200 import WideFieldModelBatch
    rng = np.random.default_rng(77)
    alpha = rng.normal(0, 0.3, (B, 10))
    beta = np.tile(np.array([1.0, 0.1, -0.05, 0.02][:nModulus]), (B, 1)) + rng.normal(0, 0.02, (B, nModulus))
    defoc = np.stack([[P["ni"] / P["lam"] * (1 + 0.01 * b), 2e4 * b, -1e4 * b] for b in range(B)])
    m = WideFieldModelBatch((N, N, Nz), B, 10, nModulus, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single,
                            lib=lib, basis=lambda nz: o.compute_zernike(nz, N, N, P["NA"], P["lam"], P["dxy"]))
    m.setPhaseBatch(alpha)
    m.setModulusBatch(beta)
    m.setDefocusBatch(defoc)
    refs = []
    for b in range(B):
        r = o.WideFieldModelOracle((N, N, Nz), 10, nModulus, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], single=single)
        r.setPhase(alpha[b]); r.setModulus(beta[b]); r.setDefocus(defoc[b])
        refs.append(r)
    return m, refs


@pytest.mark.parametrize("N,Nz,B,single", [(32, 5, 3, False), (64, 17, 2, False), (32, 4, 3, True), (256, 3, 2, False)])
def test_batch_handle_matches_independent_oracle_models(lib, N, Nz, B, single):
    m, refs = _batch_case(lib, N, Nz, B, single)
    t = tol(single)
    tj = 20 * t if single else t
    rho, phi, psi, mask = m.getRho(), m.getPhi(), m.getPsi(), m.getMaskPupil()
    psf, cpx = m.getPsf(), m.get_cpxPsf()
    q = np.stack([o.synthetic_q(N, N, Nz, seed=42 + b, single=single) for b in range(B)])
    d, p, mo = m.applyJacobianBatch(q)
    for b, r in enumerate(refs):
        np.testing.assert_array_equal(rho[b], r.rho.ravel())           # setters stay bit exact per model
        np.testing.assert_array_equal(phi[b], r.phi.ravel())
        np.testing.assert_array_equal(psi[b], r.psi.ravel())
        np.testing.assert_array_equal(mask[b], r.maskPupil.ravel())
        assert o.rel_l2(psf[b], r.getPsf()) <= t
        assert o.rel_l2(cpx[b], r.get_cpxPsf()) <= t
        assert o.rel_l2(d[b], r.apply_J_defocus(q[b])) <= tj
        assert o.rel_l2(p[b], r.apply_J_phase(q[b])) <= tj
        assert o.rel_l2(mo[b], r.apply_J_modulus(q[b])) <= tj
    # scalar setters broadcast; single-model outputs are refused on a batch handle
    m.setPhase(np.zeros(10))
    assert np.all(m.getPhi() == 0)
    with pytest.raises(RuntimeError):
        m.apply_J_phase(q)
    m.close()


def test_batch_handle_modes_and_subsets(lib):
    """Batch handle corner cases: Jacobian subsets (unselected kinds come back as zeros), the quirk-Q1 last-plane
    modulus mode per model, a chunk boundary inside a model (Nz = 19 > 16 planes per reduce chunk), nPhase = 0."""
    N, Nz, B = 32, 19, 2
    m, refs = _batch_case(lib, N, Nz, B, False)
    q = np.stack([o.synthetic_q(N, N, Nz, seed=7 + b) for b in range(B)])
    d, p, mo = m.applyJacobianBatch(q, kinds=capi.WFM_J_PHASE)
    assert np.all(d == 0) and np.all(mo == 0)
    for b, r in enumerate(refs):
        assert o.rel_l2(p[b], r.apply_J_phase(q[b])) <= 1e-12
    d, p, mo = m.applyJacobianBatch(q, kinds=capi.WFM_J_DEFOCUS | capi.WFM_J_MODULUS)
    assert np.all(p == 0)
    m.setModulusMode(True)
    d2, p2, mo2 = m.applyJacobianBatch(q, kinds=capi.WFM_J_MODULUS)
    for b, r in enumerate(refs):
        assert o.rel_l2(d[b], r.apply_J_defocus(q[b])) <= 1e-12
        assert o.rel_l2(mo[b], r.apply_J_modulus(q[b])) <= 1e-12
        r.modulus_mode = o.MODULUS_REFERENCE_LAST_PLANE
        assert o.rel_l2(mo2[b], r.apply_J_modulus(q[b])) <= 1e-12
    # errors: a table with the wrong number of rows, a defocus vector of length 2 (quirk Q4)
    with pytest.raises(ValueError):
        m.setPhaseBatch(np.zeros((B + 1, 10)))
    with pytest.raises(ValueError):
        m.setDefocusBatch(np.zeros((B, 2)))
    assert lib.wfm_batch_size(m.handle) == B
    m.close()
    hb = C.c_void_p()                                               # more models than the reduction grid can index
    assert lib.wfm_create_batch(C.byref(hb), 32, 32, 4, 70000, 1e-7, 1e-7, capi.WFM_F64, 0) == capi.WFM_ERR_INVALID_ARG


@pytest.mark.parametrize("seed", range(8))
def test_randomised_configurations_match_oracle(lib, seed):
    """Seeded sweep over shapes and parameter sets the fixed cases do not visit: odd / prime Nz (ragged last reduce
    chunk, Nz = 1), empty phase space, 1 or 4 modulus modes, off-axis defocus, fp32, a random z-slab of the stack."""
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.choice([32, 64]))
    Nz = int(rng.choice([1, 2, 5, 16, 17, 23, 33]))
    nPhase = int(rng.choice([0, 3, 10]))
    nMod = int(rng.choice([1, 4]))
    single = bool(rng.integers(0, 2))
    delta = (float(rng.normal(0, 2e4)), float(rng.normal(0, 2e4))) if rng.integers(0, 2) else None
    z0 = int(rng.integers(0, Nz))
    nzl = int(rng.integers(1, Nz - z0 + 1))
    ref, m = make_pair(N, Nz, lib, nPhase=nPhase, nModulus=nMod, single=single, delta=delta, z0=z0, nz_local=nzl)
    if nPhase:
        a = rng.normal(0, 0.3, nPhase)
        ref.setPhase(a); m.setPhase(a)
    t = tol(single)
    tj = 20 * t if single else t
    sl = slice(z0, z0 + nzl)
    assert o.rel_l2(m.getPsf(), ref.getPsf()[sl]) <= t
    assert o.rel_l2(m.get_cpxPsf(), ref.get_cpxPsf()[sl]) <= t
    # a slab's gradients are partial sums: compare with the oracle applied to q that is zero outside the slab
    q = o.synthetic_q(N, N, Nz, seed=seed, single=single)
    qz = np.zeros_like(q); qz[sl] = q[sl]
    d, p, mo = m.apply_J_all(q[sl])
    want_d, want_m = ref.apply_J_defocus(qz), ref.apply_J_modulus(qz)
    assert o.rel_l2(d, want_d) <= tj
    assert o.rel_l2(mo, want_m) <= tj
    if nPhase:
        assert o.rel_l2(p, ref.apply_J_phase(qz)) <= tj
    m.close()


def test_gpu_case_helpers_at_small_sizes(lib):
    """The bodies of the full-stack / escape-hatch GPU tests, run on the emulated build at sizes it finishes in
    seconds (a ragged slab across the z wrap, more planes than the intermediate ring holds)."""
    from tests.test_gpu_parity import chunked_host_case, escape_hatch_case, full_stack_case
    chunked_host_case(lib, 32, 40, 7, 29, False, {"WFM_HOST_CHUNKS": "4", "WFM_HOST_CHUNK_MIN_BYTES": "1"})
    chunked_host_case(lib, 32, 12, 0, 12, True, {"WFM_HOST_CHUNKS": "3", "WFM_HOST_CHUNK_MIN_BYTES": "1"})
    staged = {"WFM_FORCE_STAGED": "1", "WFM_STAGED_MIN_BYTES": "1"}          # the pageable-array path: staged by host threads
    chunked_host_case(lib, 32, 40, 7, 29, False, dict(staged, WFM_HOST_THREADS="3", WFM_HOST_SHARE_BYTES="5000",
                                                      WFM_HOST_CHUNKS="4", WFM_HOST_CHUNK_MIN_BYTES="1"))
    chunked_host_case(lib, 32, 12, 0, 12, True, dict(staged, WFM_HOST_THREADS="1", WFM_HOST_SHARE_BYTES="3000"))
    full_stack_case(lib, 32, 96, 40, 50, False)
    full_stack_case(lib, 32, 16, 3, 9, True)
    escape_hatch_case(lib, ((32, 5, 0.2), (32, 3, 0.4)))


def test_tight_ring_on_the_emulator(lib, monkeypatch):
    """WFM_PIPE_LAG=1: a ring of four planes, every producer waits for the slot's previous tenant (queue order and
    counter protocol at the least slack; the GPU twin also exercises the L2 discards)."""
    from tests.test_gpu_parity import full_stack_case
    monkeypatch.setenv("WFM_PIPE_LAG", "1")
    full_stack_case(lib, 32, 24, 2, 20, False)


@pytest.mark.parametrize("N,Nz,single", [(48, 3, False), (100, 2, False), (30, 4, True), (36, 2, False)])
def test_any_n_path_matches_oracle(lib, N, Nz, single):
    """Nx = Ny that is not a power of two in [32, 2048] (the reference's JTransforms takes any N, WFM:319): the
    any-N kernels of wfm_generic.cuh against the oracle, all three Jacobians, both modulus modes, a z-slab."""
    ref, m = make_pair(N, Nz, lib, single=single)
    t = tol(single)
    tj = 20 * t if single else t
    np.testing.assert_array_equal(m.getRho(), ref.rho.ravel())
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= t
    assert o.rel_l2(m.get_cpxPsf(), ref.get_cpxPsf()) <= t
    q = o.synthetic_q(N, N, Nz, single=single)
    d, p, mo = m.apply_J_all(q)
    want = np.concatenate([ref.apply_J_defocus(q), ref.apply_J_phase(q), ref.apply_J_modulus(q)])
    assert o.rel_l2(np.concatenate([d, p, mo]), want) <= tj
    assert o.rel_l2(m.apply_J_phase(q).data, ref.apply_J_phase(q)) <= tj
    m.setModulusMode(True)
    ref.modulus_mode = o.MODULUS_REFERENCE_LAST_PLANE
    assert o.rel_l2(m.apply_J_modulus(q).data, ref.apply_J_modulus(q)) <= tj
    m.close()
    if Nz >= 3:                                                         # z-slab of the any-N path (shard invariance)
        _, part = make_pair(N, Nz, lib, single=single, z0=1, nz_local=Nz - 1)
        assert o.rel_l2(part.getPsf(), ref.getPsf()[1:]) <= t
        part.close()
