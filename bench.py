#!/usr/bin/env python
"""Benchmark of the widefield PSF + Jacobian path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One step = setPhase(alpha) -> computePsf() -> apply_J_phase(q) over one z-slab of
512 x 512 x 256 fp64 per GPU (one unit of work = one z-plane through computePsf
and one Jacobian application).  N > 1 is z-slab weak scaling: every rank owns 256 planes of a
256*N-plane global stack and the only collective is the NCCL allreduce of the gradient vector.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "psf_plus_jacobian_z_planes_per_s_512x512_fp64"
UNIT = "z-planes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nxy", "--n", dest="n", type=int, default=512, help="Nx = Ny (use --nxy under torchrun)")
    ap.add_argument("--nz", type=int, default=256, help="z-planes per GPU")
    ap.add_argument("--single", action="store_true", help="optional fp32 mode (not the headline)")
    ap.add_argument("--kinds", type=int, default=2, help="Jacobian bits: 1 defocus, 2 phase, 4 modulus")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--cpu-planes", type=int, default=0, help="planes of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval-fg", action="store_true", help="skip the config-3 inner-loop timing (wfm_eval_fg)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# clocks (recipe in B200_PROFILING.md): sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, cmax, util = float(f[0]), float(f[1]), float(f[2])
            except ValueError:
                continue
            mx.append(cmax)
            if util > 0:
                sm.append(clk)
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores, one task per plane like WFM:287-333
# ---------------------------------------------------------------------------------------------------
def cpu_reference_planes_per_s(N, nz_global, planes, threads, repeats=1):
    """Time computePsf + apply_J_phase over `planes` z-planes with `threads` workers.  The port
    (oracle/) restates WideFieldModel's para branches; each task owns one plane (WFM:291-333,
    888-945).  Returns (planes/s, seconds)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import wfm_oracle as o
    P = o.DEFAULTS
    ref = o.WideFieldModelOracle((N, N, 2), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    ref.setPhase(o.synthetic_alpha(10))
    rho, phi, psi, mask, Z = ref.rho, ref.phi, ref.psi, ref.maskPupil, ref.Z
    nzq = min(planes, nz_global)
    q = o.synthetic_q(N, N, nz_global, nz_local=nzq)

    def task(i):
        iz = i % nzq                                                # the sample may wrap around the stack
        c, p = o.compute_psf(rho, phi, psi, nz_global, P["dz"], z0=iz, nz_local=1)
        return o.apply_J_phase(q[iz:iz + 1], c, rho, phi, psi, mask, Z, 10, nz_global, P["dz"], z0=iz)

    best = None
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(task, range(min(planes, threads))))          # warm-up (plans, page faults)
        for _ in range(repeats):
            t0 = time.perf_counter()
            g = sum(ex.map(task, range(planes)))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    assert np.all(np.isfinite(g))
    return planes / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    planes = args.cpu_planes or max(8 * cores, 64)             # ~0.25 s per step on 16 cores
    times, total = [], 0
    for _ in range(args.warmup):
        cpu_reference_planes_per_s(args.n, args.nz * args.gpus, min(planes, cores), cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, dt = cpu_reference_planes_per_s(args.n, args.nz * args.gpus, planes, cores)
        times.append(dt)
        total += planes
        if time.perf_counter() - t_all > 150:                     # keep the whole run within minutes
            break
    value = total / sum(times)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.n}x{args.n}x{args.nz} fp64 PSF+apply_J_phase per GPU (reference algorithm, host cores)",
                   "nphase": 10, "nmodulus": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{planes} planes of {args.n}x{args.n} per step, one task per plane, "
                                   f"{cores} threads (numpy/scipy restatement of WideFieldModel, not the JVM)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build_library()
    if world > 1:
        dist.barrier()
    from microtipi_b200 import WideFieldModel, _capi as capi
    from microtipi_b200.sharded import slab_bounds

    N, nzl = args.n, args.nz
    nzg = nzl * world
    z0, nz_mine = slab_bounds(nzg, world, rank)
    single = args.single
    es = 4 if single else 8
    tdt = torch.float32 if single else torch.float64
    # SURVEY.md 8d2 synthetic inputs
    P = dict(NA=1.4, lam=542e-9, ni=1.518, dxy=64.5e-9, dz=160e-9)
    m = WideFieldModel((N, N, nzg), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, single,
                       device=local, z0=z0, nz_local=nz_mine)
    lib = capi.load_library()
    # a dedicated (non-default) torch stream: the library, the NCCL allreduce and the timing events all
    # run on it (handle 0 = the legacy default stream would mean "use the handle's own stream" to the ABI)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    m.setStream(stream.cuda_stream)
    alpha = np.random.default_rng(1234).normal(0.0, 0.3, 10)
    vox = N * N * nz_mine
    dev = torch.device("cuda", local)
    q = torch.empty(vox, dtype=tdt, device=dev)
    m.fillUniform(q.data_ptr(), 42, z0 * N * N, vox)              # q resident in HBM before the timed region
    L = m.gradLength()
    grad = torch.zeros(L, dtype=torch.float64, device=dev)
    kinds = args.kinds

    x = m.parameterCoefs[m.PHASE]                                  # PSF_Estimation.java:117

    def step(i):
        x.data[:] = alpha + 1e-3 * (i % 7)                         # a new parameter vector every evaluation
        m.setParam(x)                                              # PSF_Estimation.java:202 -> setPhase -> freeMem()
        m.computePsf()
        m.applyJacobianDevice(kinds, q.data_ptr(), grad.data_ptr())
        if world > 1:
            dist.all_reduce(grad)                                  # NCCL sum of the K-vector

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    fence()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # timed region: K steps, two CUDA events on the launching stream, nothing else on it
    n0 = lib.wfm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    fence()
    ms = e0.elapsed_time(e1)
    launches = lib.wfm_launch_count() - n0
    # per-kernel durations (roofline): the same K steps again with an event pair around every kernel group
    # (the extra event records sit between the kernels, so this pass is not the one `value` is taken from)
    m.setProfiling(True)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e2.record()
    for i in range(args.steps):
        step(i)
    e3.record()
    fence()
    ms_spans = e2.elapsed_time(e3)
    ktimes = m.kernelTimes()
    m.setProfiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = nzg * args.steps / (ms * 1e-3)
    gsum = float(grad.abs().sum().item())
    if not os.environ.get("WFM_PIPE_ROLES"):                       # (single-role profiling runs compute garbage)
        assert np.isfinite(gsum) and gsum > 0.0, "gradient is not finite / zero"

    # ---- e2e: the same step through the host-buffer entry points of the C ABI -------------------------
    ne = args.e2e_steps or min(args.steps, 10)
    qbytes, pbytes = vox * es, vox * es
    hq, hp = C.c_void_p(), C.c_void_p()
    assert lib.wfm_host_alloc(C.byref(hq), qbytes) == 0 and lib.wfm_host_alloc(C.byref(hp), pbytes) == 0
    q_host = np.frombuffer((C.c_char * qbytes).from_address(hq.value), dtype=np.float32 if single else np.float64)
    q_host[:] = q.cpu().numpy()
    gout = (C.c_double * 10)()
    h = m.handle

    def e2e_step(i):
        a = np.ascontiguousarray(alpha + 1e-3 * (i % 7))
        assert lib.wfm_set_phase(h, a.ctypes.data_as(C.c_void_p), 10) == 0
        assert lib.wfm_get_psf_async(h, hp) == 0                   # computePsf + D2H of the PSF slab (2nd stream)
        assert lib.wfm_apply_j_phase(h, hq, gout, 10) == 0         # H2D of q + Jacobian + D2H of the gradient
        assert lib.wfm_wait_transfers(h) == 0                      # the PSF slab has landed in host memory
        if world > 1:
            g = torch.tensor(list(gout), dtype=torch.float64, device=dev)
            dist.all_reduce(g)
            g.cpu()

    e2e_step(0)
    fence()
    t0 = time.perf_counter()
    for i in range(ne):
        e2e_step(i)
    fence()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = nzg * ne / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    lib.wfm_host_free(hq)
    lib.wfm_host_free(hp)

    # ---- config 3: the blind-deconvolution inner loop with the data term on the device (row f1) ------
    # one evaluation = setParam(x) -> computePsf -> FFT-convolution cost + gradient -> apply_J_phase; only x
    # goes to the device, the cost and the gradient come back (PSF_Estimation.java:202-217).
    eval_fg = None
    if world == 1 and not single and not args.no_eval_fg:
        from microtipi_b200 import WeightedConvolutionCost, DoubleShapedVectorSpace
        f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, nzl), device=local)
        obj = np.zeros((nzl, N, N))
        for dz_ in range(-2, 3):
            for dy_ in range(-2, 3):
                for dx_ in range(-2, 3):
                    if dz_ * dz_ + dy_ * dy_ + dx_ * dx_ <= 6.25:
                        obj[dz_ % nzl, dy_ % N, dx_ % N] = 1.0     # solid sphere, radius 2.5 px, centred at voxel 0
        f.setPSF(obj)
        del obj
        f.setData(np.random.default_rng(7).random((nzl, N, N)) * 1e-6)
        nf = max(3, min(args.steps, 10))
        f.evalFG(m, m.PHASE, alpha)
        fence()
        t0 = time.perf_counter()
        for i in range(nf):
            cost, gfg = f.evalFG(m, m.PHASE, alpha + 1e-3 * (i % 7))
        fg_ms = (time.perf_counter() - t0) * 1e3 / nf
        assert np.isfinite(cost) and np.all(np.isfinite(gfg))
        eval_fg = {"ms_per_eval": fg_ms, "value": nzg / (fg_ms * 1e-3), "unit": UNIT, "steps": nf,
                   "h2d_bytes_per_step": 80, "d2h_bytes_per_step": 88,
                   "note": "wfm_eval_fg: PSF + 3-D FFT convolution cost/gradient (10 volume sweeps) + Jacobian, host wall clock"}
        f.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        npix = N * N
        # algorithmic bytes: k_psf_pipeline writes conj(a)+psf = 3*s*Npix per plane;
        # k_jac_pipeline reads conj(a)+q = 3*s*Npix per plane.  One launch processes the whole slab.
        alg = {"psf_pipeline": 3 * es * npix * nz_mine, "jac_pipeline": 3 * es * npix * nz_mine}
        per = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in ktimes.items()}
        dom = max(alg, key=lambda k: per.get(k, 0.0))
        dom_ms = per[dom]
        achieved = alg[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        step_bytes = 6 * es * npix * nz_mine
        step_ms = ms / args.steps
        step_gbs = step_bytes / (step_ms * 1e-3) / 1e9
        traffic = None
        tj = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tj):
            try:
                traffic = json.load(open(tj)).get(f"k_{dom}", {}).get(f"{N}x{nz_mine}x{'f32' if single else 'f64'}")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC if (N == 512 and not single) else f"psf_plus_jacobian_z_planes_per_s_{N}x{N}_{'fp32' if single else 'fp64'}",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if single else "f64", "data": "synthetic",
            "config": {"workload": f"WideFieldModel setPhase+computePsf+apply_J_phase, {N}x{N}x{nzl} "
                                   f"{'fp32' if single else 'fp64'} z-slab per GPU ({nzg} planes total)",
                       "nphase": 10, "nmodulus": 1, "NA": 1.4, "jacobian_kinds": kinds,
                       "l2_policy": f"inputs larger than L2 ({step_bytes / 1e9:.2f} GB streamed per step vs 126 MB L2)",
                       "parallelism": f"z-slab x{world}, NCCL allreduce of {L} doubles per step" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": qbytes + 80, "d2h_bytes_per_step": pbytes + 80,
                    "steps": ne, "ms_per_step": e2e_ms / ne},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "kernel": f"k_{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": dom_ms},
            "roofline_step": {"achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                              "algorithmic_bytes_per_step": step_bytes,
                              "note": "whole step: 6*s*Npix bytes per plane over the step time (per GPU)"},
            "kernel_ms_per_step": {k: round(v, 5) for k, v in per.items()},
            "kernel_timing": {"how": "second pass of the same K steps with a CUDA-event pair around every kernel group "
                                     "on the launching stream", "ms_per_step_with_event_pairs": ms_spans / args.steps},
        }
        if eval_fg:
            out["eval_fg"] = eval_fg
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            planes = args.cpu_planes or max(128 * cores, 512)      # a few seconds of wall time = tens of core-seconds
            v, dt = cpu_reference_planes_per_s(N, nzg, planes, cores)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{planes} of the {nzg} planes ({N}x{N} fp64), one task per plane on "
                                             f"{cores} threads, {dt:.1f} s (numpy/scipy restatement, not the JVM)"}
        print(json.dumps(out), flush=True)
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
