"""Multi-GPU correctness check (run under torchrun on >= 2 GPUs): z-slab shards + NCCL allreduce of the
gradient vector + NCCL all-gather of the PSF stack against the single-process oracle.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_nccl.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wfm_oracle as o  # noqa: E402  (checker only)
from microtipi_b200.sharded import ShardedWideFieldModel, slab_bounds  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    P = o.DEFAULTS
    N, Nz = 128, 10 if world == 4 else 9
    basis = lambda nz: o.compute_zernike(nz, N, N, P["NA"], P["lam"], P["dxy"])  # noqa: E731
    m = ShardedWideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"], device=local, basis=basis)
    alpha, beta = o.synthetic_alpha(10), [1.0, 0.1, -0.05, 0.02]
    m.setPhase(alpha)
    m.setModulus(beta)
    z0, nzl = slab_bounds(Nz, world, rank)
    q = o.synthetic_q(N, N, Nz, z0=z0, nz_local=nzl)
    d, p, mo = m.apply_J_all(q)
    t0 = time.perf_counter()
    psf = m.gatherPsfDevice()
    torch.cuda.synchronize()
    t_gather = time.perf_counter() - t0
    ref = o.WideFieldModelOracle((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
    ref.setPhase(alpha)
    ref.setModulus(beta)
    qf = o.synthetic_q(N, N, Nz)
    errs = {"psf": o.rel_l2(psf.cpu().numpy(), ref.getPsf()), "phase": o.rel_l2(p, ref.apply_J_phase(qf)),
            "defocus": o.rel_l2(d, ref.apply_J_defocus(qf)), "modulus": o.rel_l2(mo, ref.apply_J_modulus(qf))}
    ok = all(v <= 1e-12 for v in errs.values())
    print(f"rank {rank}/{world} slab=({z0},{nzl}) gather {t_gather * 1e3:.2f} ms errs={ {k: f'{v:.1e}' for k, v in errs.items()} } {'OK' if ok else 'FAIL'}",
          flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
