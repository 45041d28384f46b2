"""BASELINE config 5: batched PSF parameter estimation over B independent 256x256x64 bead PSFs
(throughput mode).  One evaluation = setParam(phase) -> computePsf -> apply_J_phase for every model.
Default: every model is its own handle on its own CUDA stream.  --batch: ONE batch handle
(wfm_create_batch), all planes of all models through one pipeline launch.  Under torchrun the models are split
by index across the ranks (one batch handle per GPU, no collective on the data path) and
the time is the max over ranks.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microtipi_b200 import WideFieldModel, WideFieldModelBatch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--models", type=int, default=64)
ap.add_argument("--nxy", type=int, default=256)
ap.add_argument("--nz", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--single", action="store_true")
ap.add_argument("--one-stream", action="store_true", help="all models on one stream (no overlap)")
ap.add_argument("--batch", action="store_true", help="one batch handle instead of one handle per model")
a = ap.parse_args()
P = dict(NA=1.4, lam=542e-9, ni=1.518, dxy=64.5e-9, dz=160e-9)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist
    from microtipi_b200.sharded import slab_bounds
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total_models = a.models
    b0, a.models = slab_bounds(total_models, world, rank)     # this rank's share of the models
else:
    total_models, b0 = a.models, 0
dev = torch.device("cuda", local)
tdt = torch.float32 if a.single else torch.float64
models, qs, grads = [], [], []
shared = torch.cuda.Stream()
alphas = [np.random.default_rng(1234 + b0 + b).normal(0.0, 0.3, 10) for b in range(a.models)]
if a.batch:
    bm = WideFieldModelBatch((a.nxy, a.nxy, a.nz), a.models, 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"],
                             False, a.single, device=local)
    vox = a.nxy * a.nxy * a.nz * a.models
    bq = torch.empty(vox, dtype=tdt, device=dev)
    for b in range(a.models):
        bm.fillUniform(bq.data_ptr() + b * (vox // a.models) * bq.element_size(), 42 + b, 0, vox // a.models)
    bg = torch.zeros(a.models * bm.gradLength(), dtype=torch.float64, device=dev)
    atab = np.stack(alphas)
for b in range(0 if a.batch else a.models):
    m = WideFieldModel((a.nxy, a.nxy, a.nz), 10, 1, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], False, a.single,
                       device=local)
    if a.one_stream:
        m.setStream(shared.cuda_stream)
    vox = a.nxy * a.nxy * a.nz
    q = torch.empty(vox, dtype=tdt, device=dev)
    m.fillUniform(q.data_ptr(), 42 + b, 0, vox)
    g = torch.zeros(m.gradLength(), dtype=torch.float64, device=dev)
    models.append(m); qs.append(q); grads.append(g)


def evaluate(i):
    if a.batch:
        bm.setPhaseBatch(atab + 1e-3 * (i % 7))
        bm.computePsf()
        bm.applyJacobianDevice(2, bq.data_ptr(), bg.data_ptr())
        return
    for b, m in enumerate(models):
        x = m.parameterCoefs[m.PHASE]
        x.data[:] = alphas[b] + 1e-3 * (i % 7)
        m.setParam(x)
        m.computePsf()
        m.applyJacobianDevice(2, qs[b].data_ptr(), grads[b].data_ptr())


def fence():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for i in range(3):
    evaluate(i)
fence()
t0 = time.perf_counter()
for i in range(a.steps):
    evaluate(i)
fence()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    dist.destroy_process_group()
    if rank != 0:
        sys.exit(0)
planes = total_models * a.nz * a.steps
es = 4 if a.single else 8
gbs = planes * 6 * es * a.nxy * a.nxy / dt / 1e9
print(json.dumps({"n_gpus": world, "config": f"{total_models} x {a.nxy}x{a.nxy}x{a.nz} {'fp32' if a.single else 'fp64'}",
                  "streams": "batch handle" if a.batch else ("one" if a.one_stream else "per-model"), "z_planes_per_s": planes / dt,
                  "ms_per_evaluation_of_all_models": 1e3 * dt / a.steps, "algorithmic_GBps": gbs,
                  "roofline_frac_of_6459_per_gpu": gbs / 6459.0 / world}))
