"""Single-process multi-GPU bench through the C ABI's multi-device handle (wfm_create_multi): what a one-threaded
host such as the reference's JVM caller (PSF_Estimation.java:202-217) gets from the n GPUs of a box.

    python tools/bench_multi.py --devices 8 [--nxy 512] [--nz-per-device 256] [--steps 30]

Reports (one JSON line): device-resident step (setPhase -> computePsf -> apply_J_phase with the partial K-vectors
summed on the first device over peer memory; timed with CUDA events on the first device's stream after a barrier of
all devices, max over devices) and the host-buffer e2e step (getPsf gathered + apply_J_phase with q scattered, one
PCIe link per device in parallel), each against the same handle restricted to ONE device."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microtipi_b200 import WideFieldModel, _capi as capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=0, help="0 = all visible")
ap.add_argument("--nxy", type=int, default=512)
ap.add_argument("--nz-per-device", type=int, default=256)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--e2e-steps", type=int, default=5)
ap.add_argument("--eval-fg", action="store_true", help="also time wfm_eval_fg with the z-sharded data term")
ap.add_argument("--strong", action="store_true", help="keep the TOTAL stack at --nz-per-device planes (strong scaling)")
a = ap.parse_args()
P = dict(NA=1.4, lam=542e-9, ni=1.518, dxy=64.5e-9, dz=160e-9)
lib = capi.load_library()
ndev_all = torch.cuda.device_count()
ndev = a.devices or ndev_all


def run(n_dev):
    N, nz = a.nxy, (a.nz_per_device if a.strong else a.nz_per_device * n_dev)
    m = WideFieldModel((N, N, nz), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, False,
                       devices=list(range(n_dev)))
    parts = m.parts()
    qs = []
    for (d, z0, nzl, child) in parts:
        with torch.cuda.device(d):
            q = torch.empty(N * N * nzl, dtype=torch.float64, device=f"cuda:{d}")
            assert lib.wfm_fill_uniform(child, C.c_void_p(q.data_ptr()), capi.WFM_F64, 42, z0 * N * N, q.numel()) == 0
            qs.append(q)
    grad = torch.zeros(m.gradLength(), dtype=torch.float64, device="cuda:0")
    alpha = np.random.default_rng(1234).normal(0.0, 0.3, 10)
    qptrs = [q.data_ptr() for q in qs]

    x = m.parameterCoefs[m.PHASE]                                    # PSF_Estimation.java:117

    def step(i):
        x.data[:] = alpha + 1e-3 * (i % 7)
        m.setParam(x)                                                  # PSF_Estimation.java:202 (no basis rebuild)
        m.computePsf()
        m.applyJacobianDeviceMulti(2, qptrs, grad.data_ptr())

    def sync_all():
        m.synchronize()
        for d in range(n_dev):
            torch.cuda.synchronize(d)

    for i in range(5):
        step(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(a.steps):
        step(i)
    sync_all()
    dev_ms = (time.perf_counter() - t0) * 1e3 / a.steps          # host clock around a fully drained region (all devices)
    g = grad.cpu().numpy()
    assert np.all(np.isfinite(g)) and np.abs(g).sum() > 0
    # host-buffer e2e: pinned staging arrays of the whole stack
    vox = N * N * nz
    hq, hp = C.c_void_p(), C.c_void_p()
    assert lib.wfm_host_alloc(C.byref(hq), vox * 8) == 0 and lib.wfm_host_alloc(C.byref(hp), vox * 8) == 0
    qh = np.frombuffer((C.c_char * (vox * 8)).from_address(hq.value), dtype=np.float64)
    off = 0
    for q in qs:
        qh[off:off + q.numel()] = q.cpu().numpy()
        off += q.numel()
    gout = (C.c_double * 10)()
    h = m.handle

    def e2e(i):
        al = np.ascontiguousarray(alpha + 1e-3 * (i % 7))
        assert lib.wfm_set_phase(h, al.ctypes.data_as(C.c_void_p), 10) == 0
        assert lib.wfm_get_psf_async(h, hp) == 0
        assert lib.wfm_apply_j_phase(h, hq, gout, 10) == 0
        assert lib.wfm_wait_transfers(h) == 0
    e2e(0)
    sync_all()
    t0 = time.perf_counter()
    for i in range(a.e2e_steps):
        e2e(i)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / a.e2e_steps
    lib.wfm_host_free(hq); lib.wfm_host_free(hp)
    out = {"n_dev": n_dev, "planes": nz, "device_ms_per_step": dev_ms, "device_planes_per_s": nz / (dev_ms * 1e-3),
           "e2e_ms_per_step": e2e_ms, "e2e_planes_per_s": nz / (e2e_ms * 1e-3), "e2e_bytes_per_step": 2 * vox * 8}
    # config 3 inner loop on the same devices: PSF + z-sharded FFT-convolution data term + Jacobian, one call
    if a.eval_fg and nz in (32, 64, 128, 256, 512, 1024, 2048):
        from microtipi_b200 import WeightedConvolutionCost, DoubleShapedVectorSpace
        f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, nz), devices=list(range(n_dev)))
        obj = np.zeros((nz, N, N))
        obj[:3, :3, :3] = 1.0
        f.setPSF(obj)
        del obj
        f.setData(np.random.default_rng(7).random((nz, N, N)) * 1e-6)
        f.evalFG(m, m.PHASE, alpha)
        t0 = time.perf_counter()
        for i in range(a.e2e_steps):
            cost, g = f.evalFG(m, m.PHASE, alpha + 1e-3 * (i % 7))
        fg_ms = (time.perf_counter() - t0) * 1e3 / a.e2e_steps
        assert np.isfinite(cost) and np.all(np.isfinite(g))
        out["eval_fg_ms"] = fg_ms
        out["eval_fg_planes_per_s"] = nz / (fg_ms * 1e-3)
        f.close()
    m.close()
    return out


one = run(1)
res = {"tool": "bench_multi (wfm_create_multi, one host thread)", "shape": f"{a.nxy}x{a.nxy}x{a.nz_per_device} fp64 per device",
       "one_device": one}
if ndev > 1:
    many = run(ndev)
    res["all_devices"] = many
    res["device_efficiency"] = many["device_planes_per_s"] / (ndev * one["device_planes_per_s"])
    res["e2e_efficiency"] = many["e2e_planes_per_s"] / (ndev * one["e2e_planes_per_s"])
    if "eval_fg_ms" in many and "eval_fg_ms" in one:
        res["eval_fg_efficiency"] = many["eval_fg_planes_per_s"] / (ndev * one["eval_fg_planes_per_s"])
print(json.dumps(res))
