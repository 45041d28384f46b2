"""Multi-GPU behind the C ABI (wfm_create_multi) and the peer-memory gradient exchange (wfm_exchange_*), on the
CPU-emulated build: WFM_EMU_DEVICES "devices" share the host heap, so the HOST logic (slab split, broadcast setters,
scatter / gather of the caller's arrays, partial-vector sums in device order) and the exchange PROTOCOL of
k_jac_final (slots, flags, epochs, double buffering) run exactly as compiled for the GPU.  The same cases run on real
devices in tests/test_gpu_parity.py (-m gpu)."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from oracle import wfm_oracle as o
from microtipi_b200 import WideFieldModel, _capi as capi
from tests.util import BETA4, P, emu_lib, oracle_basis


@pytest.fixture(scope="module")
def lib():
    os.environ["WFM_EMU_DEVICES"] = "4"
    yield emu_lib()
    os.environ.pop("WFM_EMU_DEVICES", None)


def multi_case(lib, N, Nz, devices, single=False):
    """Body shared with the GPU test: a multi-device model against the oracle through the reference-facing calls."""
    ref = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], single=single)
    m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single, lib=lib,
                       basis=oracle_basis(N), devices=devices)
    t = 1e-5 if single else 1e-12
    tj = 20 * t if single else t
    parts = m.parts()
    assert [p[0] for p in parts] == list(devices)
    assert [p[1] for p in parts] == list(np.cumsum([0] + [p[2] for p in parts[:-1]]))     # contiguous slabs
    assert sum(p[2] for p in parts) == Nz and max(p[2] for p in parts) - min(p[2] for p in parts) <= 1
    alpha = o.synthetic_alpha(10)
    delta = [P["ni"] / P["lam"], 1e4, -2e4]
    for mm in (ref, m):
        mm.setPhase(alpha)
        mm.setModulus(BETA4)
        mm.setDefocus(delta)
    assert m.PState == 0
    np.testing.assert_array_equal(m.getRho(), ref.rho.ravel())
    np.testing.assert_array_equal(m.getPsi(), ref.psi.ravel())
    assert o.rel_l2(m.getPsf(), ref.getPsf()) <= t                   # every slab at its offset of the caller's array
    assert o.rel_l2(m.get_cpxPsf(), ref.get_cpxPsf()) <= t
    q = o.synthetic_q(N, N, Nz, single=single)
    assert o.rel_l2(m.apply_J_phase(q).data, ref.apply_J_phase(q)) <= tj
    assert o.rel_l2(m.apply_J_defocus(q).data, ref.apply_J_defocus(q)) <= tj
    assert o.rel_l2(m.apply_J_modulus(q).data, ref.apply_J_modulus(q)) <= tj
    d, p, mo = m.apply_J_all(q)
    want = np.concatenate([ref.apply_J_defocus(q), ref.apply_J_phase(q), ref.apply_J_modulus(q)])
    assert o.rel_l2(np.concatenate([d, p, mo]), want) <= tj
    # dirty -> recompute on every device (Q5), and the last-plane modulus mode reaches the device that owns Nz-1
    a2 = alpha * 0.5
    ref.setPhase(a2); m.setPhase(a2)
    assert o.rel_l2(m.apply_J_phase(q).data, ref.apply_J_phase(q)) <= tj
    m.setModulusMode(True)
    ref.modulus_mode = o.MODULUS_REFERENCE_LAST_PLANE
    assert o.rel_l2(m.apply_J_modulus(q).data, ref.apply_J_modulus(q)) <= tj
    return ref, m, q


@pytest.mark.parametrize("N,Nz,devices,single", [(32, 7, [0, 1], False), (32, 9, [2, 0, 3, 1], False), (32, 5, [1], False),
                                                 (32, 6, [0, 1, 2], True)])
def test_multi_handle_matches_oracle(lib, N, Nz, devices, single):
    ref, m, q = multi_case(lib, N, Nz, devices, single)
    m.close()


def test_multi_handle_with_chunked_and_staged_host_paths(lib, monkeypatch):
    """The children of a multi-device handle move their slabs in plane chunks and, for pageable arrays, through their own
    staging threads (one host thread per device, each with its workers): same results through the same calls."""
    monkeypatch.setenv("WFM_HOST_CHUNKS", "2")
    monkeypatch.setenv("WFM_HOST_CHUNK_MIN_BYTES", "1")
    ref, m, q = multi_case(lib, 32, 19, [0, 1, 2], False)
    m.close()
    monkeypatch.setenv("WFM_FORCE_STAGED", "1")
    monkeypatch.setenv("WFM_STAGED_MIN_BYTES", "1")
    monkeypatch.setenv("WFM_HOST_THREADS", "2")
    monkeypatch.setenv("WFM_HOST_SHARE_BYTES", "4000")
    ref, m, q = multi_case(lib, 32, 19, [1, 0], False)
    m.close()
    ref, m, q = multi_case(lib, 32, 9, [0, 1, 2], True)
    m.close()


def test_multi_handle_device_resident_path(lib):
    """wfm_multi_apply_jacobian_dev: per-device q slabs, partial vectors land in slots on the first device."""
    N, Nz = 32, 7
    ref, m, q = multi_case(lib, N, Nz, [0, 1, 2])
    m.setModulusMode(False)
    ref.modulus_mode = o.MODULUS_INTENDED
    slabs = [np.ascontiguousarray(q[z0:z0 + n]) for (_, z0, n, _) in m.parts()]      # emulated device memory = host memory
    grad = np.zeros(m.gradLength())
    m.applyJacobianDeviceMulti(7, [s.ctypes.data for s in slabs], grad.ctypes.data)
    m.synchronize()
    want = np.concatenate([ref.apply_J_defocus(q), ref.apply_J_phase(q), ref.apply_J_modulus(q)])
    assert o.rel_l2(grad, want) <= 1e-12
    # children are reachable for device-resident use
    child = m.parts()[1][3]
    ptr = C.c_void_p()
    assert lib.wfm_device_psf(child, C.byref(ptr)) == 0 and ptr.value
    m.close()


def test_multi_handle_errors(lib):
    h = C.c_void_p()
    two = (C.c_int * 2)(0, 0)
    assert lib.wfm_create_multi(C.byref(h), 32, 32, 8, 1e-7, 1e-7, 0, two, 2) == capi.WFM_ERR_INVALID_ARG
    assert b"twice" in lib.wfm_last_error(None)
    many = (C.c_int * 4)(0, 1, 2, 3)
    assert lib.wfm_create_multi(C.byref(h), 32, 32, 3, 1e-7, 1e-7, 0, many, 4) == capi.WFM_ERR_INVALID_ARG
    bad = (C.c_int * 2)(0, 9)
    assert lib.wfm_create_multi(C.byref(h), 32, 32, 8, 1e-7, 1e-7, 0, bad, 2) != 0
    with pytest.raises(ValueError, match="Nx should equal Ny"):
        WideFieldModel((32, 64, 8), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], lib=lib, devices=[0, 1])
    m = WideFieldModel((32, 32, 8), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], lib=lib, devices=[0, 1],
                       basis=oracle_basis(32))
    with pytest.raises(RuntimeError, match="multi-device"):
        m.getPsfRolled()
    with pytest.raises(RuntimeError, match="multi-device"):
        m.setStream(0)
    with pytest.raises(ValueError):
        m.setDefocus([1.0, 2.0])                                    # Q4 is enforced on the parent too
    assert lib.wfm_multi_parts(m.handle) == 2
    m.close()


def test_peer_memory_gradient_exchange_between_ranks(lib):
    """Two "ranks" (threads; one z-slab handle each) exchange their partial K-vectors inside k_jac_final: both get the
    full gradient, bit-identical, over several calls (epochs alternate the two halves of the landing buffers)."""
    N, Nz, world = 32, 9, 3
    ref = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    bounds = [(0, 3), (3, 3), (6, 3)]
    models = [WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, False, lib=lib,
                             basis=oracle_basis(N), z0=z0, nz_local=n, device=r) for r, (z0, n) in enumerate(bounds)]
    handles = [m.exchangeExport(world) for m in models]
    for r, m in enumerate(models):
        m.exchangeConnect(r, world, handles)
        assert m.exchangeStatus() == 1
    q = o.synthetic_q(N, N, Nz)
    for it in range(5):                                             # 5 epochs: both buffer halves re-used
        alpha = o.synthetic_alpha(10) * (1.0 + 0.1 * it)
        ref.setPhase(alpha); ref.setModulus(BETA4)
        for m in models:
            m.setPhase(alpha); m.setModulus(BETA4)
        kinds = 7 if it % 2 == 0 else 2
        grads = [np.zeros(m.gradLength()) for m in models]
        slabs = [np.ascontiguousarray(q[z0:z0 + n]) for z0, n in bounds]

        def run(r):
            models[r].applyJacobianDevice(kinds, slabs[r].ctypes.data, grads[r].ctypes.data)
            models[r].synchronize()
        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        want = np.concatenate([ref.apply_J_defocus(q), ref.apply_J_phase(q), ref.apply_J_modulus(q)])
        if kinds == 2:
            want[:3] = 0.0
            want[13:] = 0.0
        for g in grads:
            assert o.rel_l2(g, want) <= 1e-12
            np.testing.assert_array_equal(g, grads[0])              # fixed rank order: identical bits on every rank
    # the host-buffer entry points ride the same exchange
    outs = [None] * world

    def run_host(r):
        z0, n = bounds[r]
        outs[r] = models[r].apply_J_phase(q[z0:z0 + n]).data
    th = [threading.Thread(target=run_host, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for g in outs:
        assert o.rel_l2(g, ref.apply_J_phase(q)) <= 1e-12
    for m in models:
        assert m.exchangeStatus() == 1
    for m in models:
        m.exchangeClose()
    for m in models:
        m.close()


def conv_multi_case(lib, N, Nz, devices):
    """Body shared with the GPU test: the z-sharded data term against the oracle, and the one-call inner loop
    (wfm_eval_fg on a multi model + multi data term) against the explicit oracle chain."""
    from microtipi_b200 import WeightedConvolutionCost, DoubleShapedVectorSpace
    rng = np.random.default_rng(5)
    shp = (Nz, N, N)
    obj, h, y = rng.normal(size=shp), rng.normal(size=shp), rng.normal(size=shp)
    w = rng.uniform(0.0, 2.0, size=shp)
    f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, Nz), lib=lib, devices=devices)
    assert lib.wfm_conv_parts(f.handle) == len(devices)
    f.setPSF(obj); f.setData(y); f.setWeights(w, True)
    g = np.zeros(h.size)
    c = f.computeCostAndGradient(0.7, h, g, True)
    c_ref, g_ref = o.weighted_convolution_cost(h, obj, y, w, 0.7)
    assert abs(c - c_ref) <= 1e-12 * abs(c_ref)
    assert o.rel_l2(g, g_ref) <= 1e-12
    c2 = f.computeCostAndGradient(0.7, h, g, False)                      # accumulate (clr = false)
    assert abs(c2 - c) <= 1e-14 * abs(c) and o.rel_l2(g, 2 * g_ref) <= 1e-12
    f.setWeights(None)
    c3 = f.computeCostAndGradient(1.0, h, g, True)
    c3_ref, g3_ref = o.weighted_convolution_cost(h, obj, y)
    assert abs(c3 - c3_ref) <= 1e-12 * abs(c3_ref) and o.rel_l2(g, g3_ref) <= 1e-12
    # the inner loop of the PSF fit on the same devices
    ref = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    m = WideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, False, lib=lib,
                       basis=oracle_basis(N), devices=devices)
    for mm in (ref, m):
        mm.setModulus(BETA4)
    for flag, x, refset, refj in (
            (m.PHASE, o.synthetic_alpha(10) + 0.01, ref.setPhase, ref.apply_J_phase),
            (m.DEFOCUS, np.array([P["ni"] / P["lam"], 1e4, -1e4]), ref.setDefocus, ref.apply_J_defocus),
            (m.MODULUS, np.array([1.0, 0.12, -0.04, 0.03]), ref.setModulus, ref.apply_J_modulus)):
        cost, gx = f.evalFG(m, flag, x)
        refset(x)
        c_ref, q_ref = o.weighted_convolution_cost(ref.getPsf(), obj, y)
        assert abs(cost - c_ref) <= 1e-11 * abs(c_ref)
        assert o.rel_l2(gx, refj(q_ref)) <= 1e-10
    f.close(); m.close()


@pytest.mark.parametrize("N,Nz,devices", [(32, 32, [0, 1]), (32, 64, [1, 3, 0]), (64, 32, [0, 1, 2, 3]), (32, 32, [2])])
def test_z_sharded_data_term_and_inner_loop(lib, N, Nz, devices):
    conv_multi_case(lib, N, Nz, devices)


def test_z_sharded_data_term_errors(lib):
    from microtipi_b200 import WeightedConvolutionCost, DoubleShapedVectorSpace
    with pytest.raises(ValueError):
        WeightedConvolutionCost.build(DoubleShapedVectorSpace(32, 32, 32), lib=lib, devices=[0, 0])
    f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(32, 32, 32), lib=lib, devices=[0, 1])
    m1 = WideFieldModel((32, 32, 32), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], lib=lib, basis=oracle_basis(32))
    f.setPSF(np.ones((32, 32, 32))); f.setData(np.zeros((32, 32, 32)))
    with pytest.raises(ValueError, match="multi-device"):
        f.evalFG(m1, m1.PHASE, np.zeros(10))                                # single-device model, sharded data term
    m3 = WideFieldModel((32, 32, 32), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], lib=lib, basis=oracle_basis(32),
                        devices=[0, 1, 2])
    with pytest.raises(ValueError, match="different device lists"):
        f.evalFG(m3, m3.PHASE, np.zeros(10))
    f.close(); m1.close(); m3.close()
