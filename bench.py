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
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json config to time (3 = the headline 512x512x256 per GPU)")
    ap.add_argument("--nxy", "--n", dest="n", type=int, default=0, help="override Nx = Ny of the config (use --nxy under torchrun)")
    ap.add_argument("--nz", type=int, default=0, help="override the z-planes of the config (per GPU for weak configs)")
    ap.add_argument("--sustain", type=float, default=2.5, help="seconds of the sustained-roofline loop (0 = skip)")
    ap.add_argument("--no-others", dest="others", action="store_false", help="skip the extra lines of configs 2, 4, 5")
    ap.add_argument("--quick", action="store_true", help="headline numbers only (variant A/B runs)")
    ap.add_argument("--quick-sustain", action="store_true", help="with --quick: keep the sustained loop")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded 128^2 parity check (single-size experiment builds)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: how the K-vector is summed over the ranks (peer = inside k_jac_final over CUDA-IPC "
                         "peer memory; nccl = one all-reduce per step; auto = peer if every rank can map its peers)")
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
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, window=None):
        """Summary of the samples taken inside `window` = (t0, t1) on the perf_counter clock (widened by 60 ms on both
        sides, one sampling period plus nvidia-smi's own latency; all samples when no window is given)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, cmax = float(f[0]), float(f[1])
            except ValueError:
                continue
            mx.append(cmax)
            if window is not None and not (window[0] - 0.06 <= ts <= window[1] + 0.06):
                continue
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
    """The reference's CPU implementation of the path (the oracle port: there is no JVM / TiPi / JTransforms here) on
    all host cores, on the B200 arm's own `config`; each step is a bounded sample of that workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    n, planes_local, planes_global, models = resolve_shape(args, world)
    nz_model = planes_global if not models else CONFIGS[5]["nz"]
    cores = os.cpu_count() or 1
    planes = args.cpu_planes or max(8 * cores, 64)             # ~0.25 s per step on 16 cores at 512 x 512
    times, total = [], 0
    for _ in range(args.warmup):
        cpu_reference_planes_per_s(n, nz_model, min(planes, cores), cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, dt = cpu_reference_planes_per_s(n, nz_model, planes, cores)
        times.append(dt)
        total += planes
        if time.perf_counter() - t_all > 150:                     # keep the whole run within minutes
            break
    value = total / sum(times)
    jvm = jvm_probe()
    out = {
        "impl": "reference", "metric": METRIC if (n == 512 and args.config == 3 and not args.single) else
        f"psf_plus_jacobian_z_planes_per_s_{n}x{n}_fp64", "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": CONFIGS[args.config]["mode"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": describe_config(args.config, n, planes_local, planes_global, False, args.kinds, world, models),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{planes} planes of {n}x{n} per step (of the {planes_global} of the config), one task "
                                   f"per plane, {cores} threads (numpy/scipy restatement of WideFieldModel, not the JVM: "
                                   f"java={jvm['java']}, jars={len(jvm['jars'])})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "jvm": jvm,
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


# BASELINE.json configs as bench workloads.  "weak": `nz` planes (or `models` models) PER GPU; "strong": the global
# stack / batch is split across the ranks.
CONFIGS = {
    1: dict(n=64, nz=32, mode="weak", name="config 1: 64x64x32 fp64"),
    2: dict(n=256, nz=128, mode="weak", name="config 2: 256x256x128 fp64"),
    3: dict(n=512, nz=256, mode="weak", name="config 3 / headline: 512x512x256 fp64 z-slab per GPU"),
    4: dict(n=1024, nz=512, mode="strong", name="config 4: 1024x1024x512 fp64, z-slabs split across the GPUs"),
    5: dict(n=256, nz=64, models=64, mode="strong", name="config 5: 64 independent 256x256x64 models, split by model index"),
}
PHYS = dict(NA=1.4, lam=542e-9, ni=1.518, dxy=64.5e-9, dz=160e-9)     # SURVEY.md 8d2 synthetic inputs


def jvm_probe():
    """is there a JVM (and TiPi / JTransforms jars) on this box to time the real reference with?"""
    import glob
    import shutil
    out = {"java": shutil.which("java"), "javac": shutil.which("javac"), "version": None, "jars": []}
    if out["java"]:
        try:
            r = subprocess.run([out["java"], "-version"], capture_output=True, text=True, timeout=20)
            out["version"] = (r.stderr or r.stdout).strip().splitlines()[0]
        except Exception as e:                                           # noqa: BLE001
            out["version"] = f"probe failed: {e}"
    for root in ("/usr/share/java", "/opt", "/usr/local", os.path.expanduser("~")):
        for pat in ("*[jJ][tT]ransforms*.jar", "*TiPi*.jar", "*tipi*.jar"):
            try:
                out["jars"] += glob.glob(os.path.join(root, "**", pat), recursive=True)[:4]
            except Exception:                                            # noqa: BLE001
                pass
    out["usable"] = bool(out["java"] and out["javac"] and out["jars"])
    return out


def git_head():
    try:
        head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                              timeout=5).stdout.strip()
        if head:
            return head
    except Exception:                                                    # noqa: BLE001
        pass
    try:                                  # no .git on the GPU box: the commit the library was built from (__graft_entry__.build_library)
        return open(os.path.join(ROOT, "microtipi_b200", "csrc", "BUILD_COMMIT")).read().strip() or None
    except OSError:
        return None


def describe_config(cfg_id, n, planes_local, planes_global, single, kinds, world, models):
    """The `config` object of the JSON line -- identical for the B200 arm and the reference arm."""
    cfg = CONFIGS[cfg_id]
    es = 4 if single else 8
    nmod = 4 if (kinds & 4) else 1
    L = 3 + 10 + nmod
    alg_step = 6 * es * n * n * planes_local
    return {"workload": f"WideFieldModel setPhase+computePsf+apply_J_phase, {cfg['name']}: {n}x{n}x"
                        f"{planes_local} {'fp32' if single else 'fp64'} planes on each GPU ({planes_global} planes total)",
            "baseline_config": cfg_id, "nphase": 10, "nmodulus": nmod, "NA": 1.4, "jacobian_kinds": kinds,
            "l2_policy": f"inputs larger than L2 ({alg_step / 1e9:.2f} GB streamed per step vs 126 MB L2)",
            "parallelism": (f"z-slab x{world}, the K-vector of {L} doubles is summed over the ranks once per step"
                            if not models else f"models split by index x{world}, no collective") if world > 1 else "single GPU"}


def resolve_shape(args, world):
    """(n, planes per GPU, planes in total, models per GPU) of the selected config, with the --nxy / --nz overrides."""
    cfg = CONFIGS[args.config]
    n = args.n or cfg["n"]
    nz = args.nz or cfg["nz"]
    if args.config == 5:
        return n, cfg["models"] * nz // world, cfg["models"] * nz, cfg["models"] // world
    if cfg["mode"] == "weak":
        return n, nz, nz * world, 0
    return n, nz // world, nz, 0


class Workload:
    """One rank's share of a PSF + Jacobian evaluation: setParam(phase) -> computePsf -> apply_J (device resident).
    z-slab of one model (`models` == 0) or a batch of independent models (config 5)."""

    def __init__(self, torch, dist, local, world, rank, n, nzg, z0, nzl, single, kinds, stream, models=0, exchange="nccl"):
        import numpy as np
        from microtipi_b200 import WideFieldModel, WideFieldModelBatch
        self.torch, self.dist, self.world, self.np = torch, dist, world, np
        self.n, self.nzg, self.nzl, self.single, self.kinds, self.models = n, nzg, nzl, single, kinds, models
        self.es = 4 if single else 8
        tdt = torch.float32 if single else torch.float64
        dev = torch.device("cuda", local)
        nmod = 4 if (kinds & 4) else 1
        if models:
            self.m = WideFieldModelBatch((n, n, nzl), models, 10, nmod, PHYS["NA"], PHYS["lam"], PHYS["ni"], PHYS["dxy"],
                                         PHYS["dz"], False, single, device=local)
            self.planes_local = models * nzl
            self.alpha = np.stack([np.random.default_rng(1234 + z0 + b).normal(0.0, 0.3, 10) for b in range(models)])
        else:
            self.m = WideFieldModel((n, n, nzg), 10, nmod, PHYS["NA"], PHYS["lam"], PHYS["ni"], PHYS["dxy"], PHYS["dz"],
                                    False, single, device=local, z0=z0, nz_local=nzl)
            self.planes_local = nzl
            self.alpha = np.random.default_rng(1234).normal(0.0, 0.3, 10)
        self.m.setStream(stream.cuda_stream)
        self.vox = n * n * self.planes_local
        self.q = torch.empty(self.vox, dtype=tdt, device=dev)
        if models:
            per = n * n * nzl
            for b in range(models):
                self.m.fillUniform(self.q.data_ptr() + b * per * self.es, 42 + z0 + b, 0, per)
        else:
            self.m.fillUniform(self.q.data_ptr(), 42, z0 * n * n, self.vox)   # q resident in HBM before any timed region
        self.L = self.m.gradLength()
        self.grad = torch.zeros(self.L * max(models, 1), dtype=torch.float64, device=dev)
        self.x = None if models else self.m.parameterCoefs[self.m.PHASE]      # PSF_Estimation.java:117
        self.reduce = (world > 1 and not models)
        self.exchange = "none"
        if self.reduce:
            from microtipi_b200.sharded import connect_peer_exchange
            self.exchange = "nccl"
            if exchange in ("peer", "auto"):
                if connect_peer_exchange(self.m, dist):
                    self.exchange = "peer"                                        # k_jac_final sums over the ranks itself
                elif exchange == "peer":
                    raise SystemExit("bench.py: --exchange peer could not be set up on every rank")

    def step(self, i, kinds=None):
        m = self.m
        if self.models:
            m.setPhaseBatch(self.alpha + 1e-3 * (i % 7))
        else:
            self.x.data[:] = self.alpha + 1e-3 * (i % 7)                      # a new parameter vector every evaluation
            m.setParam(self.x)                                                # PSF_Estimation.java:202 -> setPhase -> freeMem()
        m.computePsf()
        m.applyJacobianDevice(self.kinds if kinds is None else kinds, self.q.data_ptr(), self.grad.data_ptr())
        if self.reduce and self.exchange == "nccl":
            self.dist.all_reduce(self.grad)                                   # NCCL sum of the K-vector

    def alg_bytes_per_step(self):
        return 6 * self.es * self.n * self.n * self.planes_local              # SURVEY 8 d3, per GPU

    def close(self):
        if self.exchange == "peer":
            self.m.synchronize()
            self.dist.barrier()                                               # nobody unmaps while a peer may still store
            self.m.exchangeStatus()
            self.m.exchangeClose()
        self.m.close()
        del self.q, self.grad


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
    from microtipi_b200 import _capi as capi
    from microtipi_b200.sharded import slab_bounds
    lib = capi.load_library()
    dev = torch.device("cuda", local)
    # a dedicated (non-default) torch stream: the library, the NCCL allreduce and the timing events all
    # run on it (handle 0 = the legacy default stream would mean "use the handle's own stream" to the ABI)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make(cfg_id, n=None, nz=None, single=False, kinds=2):
        """Workload of one BASELINE config on this rank (None when the rank has no share)."""
        cfg = dict(CONFIGS[cfg_id])
        if n:
            cfg["n"] = n
        if nz:
            cfg["nz"] = nz
        if cfg_id == 5:
            b0, nb = slab_bounds(cfg["models"], world, rank)
            w = Workload(torch, dist, local, world, rank, cfg["n"], cfg["nz"], b0, cfg["nz"], single, kinds, stream, models=nb,
                         exchange=args.exchange)
            return w, cfg["models"] * cfg["nz"], cfg
        if cfg["mode"] == "weak":
            nzg = cfg["nz"] * world
        else:
            nzg = cfg["nz"]
        z0, nzl = slab_bounds(nzg, world, rank)
        w = Workload(torch, dist, local, world, rank, cfg["n"], nzg, z0, nzl, single, kinds, stream, exchange=args.exchange)
        return w, nzg, cfg

    def timed(w, steps, warmup, kinds=None):
        """(ms for `steps` steps, max over ranks): two CUDA events on the launching stream, nothing else on it."""
        for i in range(warmup):
            w.step(i, kinds)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        e0.record()
        for i in range(steps):
            w.step(i, kinds)
        e1.record()
        fence()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- parity before timing: a small sharded run against the oracle (checker only), on every rank ----------
    parity = None if (args.quick or args.no_parity) else check_parity(torch, dist, local, world, rank, stream, args.exchange)

    cfg_id = args.config
    single = args.single
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()       # nvidia-smi needs a few hundred ms before its first sample: started before the set-up
    w, planes_global, cfg = make(cfg_id, args.n, args.nz, single, args.kinds)
    N, nzl, nzg = w.n, w.nzl, w.nzg
    es = w.es
    warm = max(args.warmup, 3)
    t_clock0 = time.perf_counter()
    for i in range(warm):
        w.step(i)
    fence()
    # ---- timed region: K steps --------------------------------------------------------------------------------
    n0 = lib.wfm_launch_count()
    ms = timed(w, args.steps, 0)
    launches = lib.wfm_launch_count() - n0
    value = planes_global * args.steps / (ms * 1e-3)
    gsum = float(w.grad.abs().sum().item())
    if not os.environ.get("WFM_PIPE_ROLES") and not os.environ.get("WFM_BENCH_NO_CHECK"):   # (timing probes compute garbage)
        assert np.isfinite(gsum) and gsum > 0.0, "gradient is not finite / zero"

    # ---- per-kernel durations (roofline): the same K steps with an event pair around every kernel group ---------
    # (the extra event records sit between the kernels, so this pass is not the one `value` is taken from)
    w.m.setProfiling(True)
    ms_spans = timed(w, args.steps, 0)
    ktimes = w.m.kernelTimes()
    w.m.setProfiling(False)

    # ---- distribution over the run (median and best): blocks of 5 steps, one event pair per block ----
    blk = 5
    nblk = max(3, args.steps // blk)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(nblk + 1)]
    fence()
    evs[0].record()
    for b in range(nblk):
        for i in range(blk):
            w.step(b * blk + i)
        evs[b + 1].record()
    fence()
    # clocks: the samples that fall between the first warm-up step and the end of this pass (same kernels throughout)
    clocks = sampler.stop((t_clock0, time.perf_counter())) if rank == 0 else None
    per_step = sorted(evs[b].elapsed_time(evs[b + 1]) / blk for b in range(nblk))
    dist_ms = {"median": max_over_ranks(per_step[len(per_step) // 2]), "best": max_over_ranks(per_step[0]),
               "worst": max_over_ranks(per_step[-1]), "blocks": nblk, "steps_per_block": blk}

    # ---- apply_J_all (kinds = 7: defocus + phase + modulus from one adjoint pass), reported additionally ------------
    j_all = None
    if not w.models and not args.quick:
        ms7 = timed(w, max(10, args.steps // 4), 2, kinds=7)
        j_all = {"ms_per_step": ms7 / max(10, args.steps // 4), "value": planes_global * max(10, args.steps // 4) / (ms7 * 1e-3),
                 "unit": UNIT, "note": "setPhase + computePsf + apply_J_all (three Jacobians, one adjoint FFT pass)"}

    # ---- sustained: the same step back to back for >= args.sustain seconds (the 1 kW power cap pulls the clock down) --
    sustained = None
    if args.sustain > 0 and (not args.quick or args.quick_sustain):
        s2 = ClockSampler(local)
        t_target = args.sustain * 1e3
        est = ms / args.steps
        nsteps = max(args.steps, int(t_target / est) + 1)
        for i in range(int(0.3 * nsteps)):                                   # reach the thermal / power steady state first
            w.step(i)
        if rank == 0:
            s2.start()
        t_s0 = time.perf_counter()
        ms_s = timed(w, nsteps, 0)
        c2 = s2.stop((t_s0 + 0.1, time.perf_counter())) if rank == 0 else None
        sustained = {"steps": nsteps, "seconds": ms_s * 1e-3, "ms_per_step": ms_s / nsteps,
                     "value": planes_global * nsteps / (ms_s * 1e-3), "clocks": c2}

    # ---- e2e: the same step through the host-buffer entry points of the C ABI ----------------------------------
    e2e = e2e_pageable = None
    if not w.models:
        e2e, e2e_pageable = run_e2e(args, torch, dist, lib, w, world, dev, fence, max_over_ranks, planes_global)

    # ---- config 3: the blind-deconvolution inner loop with the data term on the device (row f1) ------------------
    eval_fg = None
    if world == 1 and not single and not args.no_eval_fg and not w.models and not args.quick and cfg_id == 3:
        eval_fg = run_eval_fg(args, torch, w, fence, planes_global)

    per = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in ktimes.items()}
    alg_step = w.alg_bytes_per_step()
    main_close = w.close
    main = dict(N=N, nzl=nzl, nzg=nzg, planes_local=w.planes_local, L=w.L, models=w.models, exchange=w.exchange)
    main_close()

    # ---- the other BASELINE configs, a few steps each, as extra keys of the same line ------------------------------
    others = {}
    if args.others and not args.quick and cfg_id == 3 and not single:
        for oid in (2, 4, 5):
            try:
                ow, oplanes, ocfg = make(oid)
                osteps = 10 if oid != 2 else 30
                oms = timed(ow, osteps, 3)
                peak, _ = measured_peak()
                gbs = ow.alg_bytes_per_step() * osteps / (oms * 1e-3) / 1e9
                ent = {"workload": ocfg["name"], "value": oplanes * osteps / (oms * 1e-3), "unit": UNIT,
                       "ms_per_step": oms / osteps, "steps": osteps, "planes_per_gpu": ow.planes_local,
                       "roofline_step_frac_per_gpu": gbs / peak, "scaling": ocfg["mode"]}
                if oid == 4 and world > 1:                                    # the PSF all-gather, timed apart
                    ent["psf_allgather"] = time_allgather(torch, dist, ow, world, fence, max_over_ranks)
                ow.close()
                others[f"config{oid}"] = ent
            except Exception as e:                                          # noqa: BLE001
                others[f"config{oid}"] = {"error": str(e)[:200]}
            fence()

    if rank == 0:
        peak, peak_src = measured_peak()
        npix = main["N"] * main["N"]
        # algorithmic bytes: k_psf_pipeline writes conj(a)+psf = 3*s*Npix per plane;
        # k_jac_pipeline reads conj(a)+q = 3*s*Npix per plane.  One launch processes the whole slab.
        alg = {"psf_pipeline": 3 * es * npix * main["planes_local"], "jac_pipeline": 3 * es * npix * main["planes_local"]}
        dom = max(alg, key=lambda k: per.get(k, 0.0))
        dom_ms = per[dom]
        achieved = alg[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        step_ms = ms / args.steps
        step_gbs = alg_step / (step_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tj = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tj):
            try:
                tjd = json.load(open(tj))
                traffic = tjd.get(f"k_{dom}", {}).get(f"{main['N']}x{main['planes_local']}x{'f32' if single else 'f64'}")
                traffic_src = {"file": "profiles/traffic.json", "capture": tjd.get("_source"), "capture_commit": tjd.get("_commit"),
                               "bench_commit": git_head()}
            except Exception:
                traffic = None
        out = {
            "metric": METRIC if (main["N"] == 512 and not single and cfg_id == 3) else
            f"psf_plus_jacobian_z_planes_per_s_{main['N']}x{main['N']}_{'fp32' if single else 'fp64'}",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": cfg["mode"], "vs_baseline": None,
            "dtype": "f32" if single else "f64", "data": "synthetic",
            "config": describe_config(cfg_id, main["N"], main["planes_local"], planes_global, single, args.kinds, world,
                                      main["models"]),
            "exchange": {"none": None, "peer": "inside k_jac_final over NVLink peer memory (CUDA IPC landing buffers, flags, fixed "
                                               "rank order)", "nccl": "one NCCL all-reduce per step"}[main["exchange"]],
            "clocks": clocks,
            "ms_per_step_distribution": dist_ms,
            "e2e": e2e,
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "kernel": f"k_{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": dom_ms},
            "roofline_step": {"achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                              "algorithmic_bytes_per_step": alg_step,
                              "note": "whole step: 6*s*Npix bytes per plane over the step time (per GPU)"},
            "kernel_ms_per_step": {k: round(v, 5) for k, v in per.items()},
            "kernel_timing": {"how": "second pass of the same K steps with a CUDA-event pair around every kernel group "
                                     "on the launching stream", "ms_per_step_with_event_pairs": ms_spans / args.steps},
            "parity": parity,
            "jvm": jvm_probe(),
        }
        if e2e is None:
            out["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                          "note": "batch workload: no host-buffer leg"}
        if e2e_pageable:
            out["e2e_pageable"] = e2e_pageable
        if sustained:
            sg = alg_step / (sustained["ms_per_step"] * 1e-3) / 1e9
            out["roofline_sustained"] = {"achieved": sg, "peak": peak, "unit": "GB/s", "frac": sg / peak, **sustained,
                                         "note": f"the same step back to back for {sustained['seconds']:.1f} s after a 30 % "
                                                 "lead-in; sw_power_cap is expected here and is a note, not a rejection"}
        if j_all:
            j_all["roofline_step_frac"] = alg_step / (j_all["ms_per_step"] * 1e-3) / 1e9 / peak
            out["apply_j_all"] = j_all
        if eval_fg:
            out["eval_fg"] = eval_fg
        if others:
            out["other_configs"] = others
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            cores = os.cpu_count() or 1
            planes = args.cpu_planes or max(128 * cores, 512)      # a few seconds of wall time = tens of core-seconds
            v, dt = cpu_reference_planes_per_s(main["N"], nzg, planes, cores)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{planes} of the {nzg} planes ({main['N']}x{main['N']} fp64), one task per plane on "
                                             f"{cores} threads, {dt:.1f} s (numpy/scipy restatement, not the JVM: "
                                             f"{'no JVM on this box' if not out['jvm']['usable'] else 'JVM present'})"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def check_parity(torch, dist, local, world, rank, stream, exchange="auto"):
    """128 x 128 x (2*world + 1) planes, sharded like the timed run (ragged slabs), against the oracle: the PSF slab
    of every rank, and the three gradients after the same exchange the timed step uses.  rel-L2, asserted <= 1e-12."""
    import numpy as np
    from oracle import wfm_oracle as o            # checker only (never on the measured path)
    from microtipi_b200.sharded import ShardedWideFieldModel, slab_bounds
    N, Nz = 128, 2 * world + 1
    P = o.DEFAULTS
    ref = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    m = ShardedWideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], device=local,
                              basis=lambda nz: ref.Z[:nz], exchange=exchange if world > 1 else "nccl")
    alpha, beta = o.synthetic_alpha(10), [1.0, 0.1, -0.05, 0.02]
    for mm in (ref, m):
        mm.setPhase(alpha)
        mm.setModulus(beta)
    z0, nzl = slab_bounds(Nz, world, rank)
    qf = o.synthetic_q(N, N, Nz)
    dev = torch.device("cuda", local)
    q = torch.from_numpy(np.ascontiguousarray(qf[z0:z0 + nzl])).to(dev)
    g = torch.zeros(m.gradLength(), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    m.applyJacobianDeviceAllReduce(7, q, g)        # handle stream -> event -> NCCL on the torch stream
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    want = np.concatenate([ref.apply_J_defocus(qf), ref.apply_J_phase(qf), ref.apply_J_modulus(qf)])
    errs = {"psf_slab": o.rel_l2(m.getPsf(), ref.getPsf()[z0:z0 + nzl]), "grad_defocus": o.rel_l2(g[:3], want[:3]),
            "grad_phase": o.rel_l2(g[3:13], want[3:13]), "grad_modulus": o.rel_l2(g[13:], want[13:])}
    used_exchange = m.exchange
    m.close()
    worst = max(errs.values())
    if world > 1:
        t = torch.tensor([worst], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
    assert worst <= 1e-12, f"rank {rank}: sharded parity failure {errs}"
    return {"shape": f"{N}x{N}x{Nz} over {world} rank(s)", "exchange": used_exchange, "rel_l2_rank0": {k: float(f"{v:.3e}") for k, v in errs.items()},
            "worst_over_ranks": float(f"{worst:.3e}"), "tolerance": 1e-12, "checker": "oracle/wfm_oracle.py"}


def run_e2e(args, torch, dist, lib, w, world, dev, fence, max_over_ranks, planes_global):
    """The step through the reference-facing host-buffer entry points: H2D of q and D2H of the PSF slab inside the
    timed region.  Twice: pinned staging buffers (wfm_host_alloc) and ordinary pageable arrays -- the closest stand-in
    for TiPi's Java-heap double[]."""
    import numpy as np
    ne = args.e2e_steps or min(args.steps, 10)
    single = w.single
    qbytes = pbytes = w.vox * w.es
    hq, hp = C.c_void_p(), C.c_void_p()
    assert lib.wfm_host_alloc(C.byref(hq), qbytes) == 0 and lib.wfm_host_alloc(C.byref(hp), pbytes) == 0
    dt = np.float32 if single else np.float64
    q_host = np.frombuffer((C.c_char * qbytes).from_address(hq.value), dtype=dt)
    q_host[:] = w.q.cpu().numpy()
    gout = (C.c_double * 10)()
    h = w.m.handle
    alpha = w.alpha

    def finish():
        if world > 1:
            g = torch.tensor(list(gout), dtype=torch.float64, device=dev)
            dist.all_reduce(g)
            g.cpu()

    def step_pinned(i):
        a = np.ascontiguousarray(alpha + 1e-3 * (i % 7))
        assert lib.wfm_set_phase(h, a.ctypes.data_as(C.c_void_p), 10) == 0
        assert lib.wfm_get_psf_async(h, hp) == 0                   # computePsf + D2H of the PSF slab (2nd stream)
        assert lib.wfm_apply_j_phase(h, hq, gout, 10) == 0         # H2D of q + Jacobian + D2H of the gradient
        assert lib.wfm_wait_transfers(h) == 0                      # the PSF slab has landed in host memory
        finish()

    q_page = np.array(q_host, copy=True)                           # plain malloc'ed arrays
    p_page = np.empty(w.vox, dtype=dt)

    def step_pageable(i):
        a = np.ascontiguousarray(alpha + 1e-3 * (i % 7))
        assert lib.wfm_set_phase(h, a.ctypes.data_as(C.c_void_p), 10) == 0
        assert lib.wfm_get_psf(h, p_page.ctypes.data_as(C.c_void_p)) == 0
        assert lib.wfm_apply_j_phase(h, q_page.ctypes.data_as(C.c_void_p), gout, 10) == 0
        finish()

    res = []
    for fn, n_it in ((step_pinned, ne), (step_pageable, max(2, ne // 2))):
        fn(0)
        fence()
        t0 = time.perf_counter()
        for i in range(n_it):
            fn(i)
        fence()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        res.append({"value": planes_global * n_it / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": qbytes + 80,
                    "d2h_bytes_per_step": pbytes + 80, "steps": n_it, "ms_per_step": ms / n_it})
    res[0]["buffers"] = "pinned (wfm_host_alloc), PSF read-back on a second stream beside the H2D of q"
    res[1]["buffers"] = ("pageable numpy arrays (stand-in for Java-heap double[]), synchronous wfm_get_psf; the library stages "
                         "them through pinned slots with its own host threads (WFM_NO_STAGED=1: plain cudaMemcpy)")
    lib.wfm_host_free(hq)
    lib.wfm_host_free(hp)
    return res[0], res[1]


def run_eval_fg(args, torch, w, fence, planes_global):
    """One evaluation = setParam(x) -> computePsf -> FFT-convolution cost + gradient -> apply_J_phase; only x goes to
    the device, the cost and the gradient come back (PSF_Estimation.java:202-217)."""
    import numpy as np
    from microtipi_b200 import WeightedConvolutionCost, DoubleShapedVectorSpace
    N, nzl = w.n, w.nzl
    f = WeightedConvolutionCost.build(DoubleShapedVectorSpace(N, N, nzl), device=torch.cuda.current_device())
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
    f.evalFG(w.m, w.m.PHASE, w.alpha)
    fence()
    t0 = time.perf_counter()
    for i in range(nf):
        cost, gfg = f.evalFG(w.m, w.m.PHASE, w.alpha + 1e-3 * (i % 7))
    fg_ms = (time.perf_counter() - t0) * 1e3 / nf
    assert np.isfinite(cost) and np.all(np.isfinite(gfg))
    f.close()
    return {"ms_per_eval": fg_ms, "value": planes_global / (fg_ms * 1e-3), "unit": UNIT, "steps": nf,
            "h2d_bytes_per_step": 80, "d2h_bytes_per_step": 88,
            "note": "wfm_eval_fg: PSF + 3-D FFT convolution cost/gradient (10 volume sweeps) + Jacobian, host wall clock"}


def time_allgather(torch, dist, w, world, fence, max_over_ranks):
    """NCCL all-gather of the PSF slabs (8*Npix*Nz bytes on every rank): off the hot path, timed apart."""
    local = w.m.devicePsfTensor()
    w.m.synchronize()
    out = torch.empty((w.nzg, w.n, w.n), dtype=local.dtype, device=local.device)
    best = None
    for _ in range(3):
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather_into_tensor(out, local.contiguous())
        e1.record()
        fence()
        t = max_over_ranks(e0.elapsed_time(e1))
        best = t if best is None else min(best, t)
    nbytes = out.numel() * out.element_size()
    return {"ms": best, "bytes_on_every_rank": nbytes, "algbw_GBps": nbytes / (best * 1e-3) / 1e9,
            "received_GBps_per_gpu": nbytes * (world - 1) / world / (best * 1e-3) / 1e9}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
