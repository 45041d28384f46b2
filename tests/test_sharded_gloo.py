"""world_size-2 gloo test of the z-slab sharding host logic, on the CPU-emulated
build of the library: each rank owns a slab, gradients are summed with ONE allreduce."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, Nz, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["WFM_EMU_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import wfm_oracle as o
    from microtipi_b200.sharded import ShardedWideFieldModel, slab_bounds
    from tests.util import P, BETA4, emu_lib, oracle_basis
    m = ShardedWideFieldModel((N, N, Nz), 10, 4, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"],
                              lib=emu_lib(), basis=oracle_basis(N))
    m.setPhase(o.synthetic_alpha(10))
    m.setModulus(BETA4)
    z0, nzl = slab_bounds(Nz, world, rank)
    assert (m.z0, m.nz_local) == (z0, nzl)
    q = o.synthetic_q(N, N, Nz, z0=z0, nz_local=nzl)      # shard reproducible from the global index
    d, p, mo = m.apply_J_all(q)
    gp = m.apply_J_phase(q)
    psf = m.getPsf(gather=True)
    if rank == 0:
        np.savez(out, d=d, p=p, mo=mo, gp=gp, psf=psf)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_model(tmp_path):
    from oracle import wfm_oracle as o
    from microtipi_b200.sharded import slab_bounds
    assert [slab_bounds(7, 2, r) for r in range(2)] == [(0, 4), (4, 3)]
    assert [slab_bounds(512, 8, r) for r in range(8)] == [(64 * r, 64) for r in range(8)]
    N, Nz = 32, 7
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, N, Nz, out), nprocs=2, join=True)
    got = np.load(out)
    from tests.util import make_pair
    ref, _m = make_pair(N, Nz, __import__("tests.util", fromlist=["emu_lib"]).emu_lib())
    _m.close()
    q = o.synthetic_q(N, N, Nz)
    assert o.rel_l2(got["psf"], ref.getPsf()) <= 1e-12
    assert o.rel_l2(got["p"], ref.apply_J_phase(q)) <= 1e-12
    assert o.rel_l2(got["gp"], ref.apply_J_phase(q)) <= 1e-12
    assert o.rel_l2(got["d"], ref.apply_J_defocus(q)) <= 1e-12
    assert o.rel_l2(got["mo"], ref.apply_J_modulus(q)) <= 1e-12


def _batch_worker(rank, world, port, N, Nz, B, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["WFM_EMU_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import wfm_oracle as o
    from microtipi_b200.sharded import ShardedWideFieldModelBatch
    from tests.util import P, emu_lib, oracle_basis
    rng = np.random.default_rng(5)
    alpha = rng.normal(0, 0.3, (B, 10))                       # the GLOBAL tables, identical on every rank
    beta = np.tile([1.0, 0.1], (B, 1)) + rng.normal(0, 0.02, (B, 2))
    m = ShardedWideFieldModelBatch((N, N, Nz), B, 10, 2, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"],
                                   lib=emu_lib(), basis=oracle_basis(N))
    m.setPhaseBatch(alpha)
    m.setModulusBatch(beta)
    q = np.stack([o.synthetic_q(N, N, Nz, seed=42 + b) for b in range(m.b0, m.b0 + m.nb_local)])
    d, p, mo = m.applyJacobianBatch(q, gather=True)
    if rank == 0:
        np.savez(out, d=d, p=p, mo=mo, alpha=alpha, beta=beta, shares=np.array([m.b0, m.nb_local]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_batch_sharded_by_model(tmp_path):
    """Config 5 across ranks: models split by index, no data-path collective; gathered gradient rows == per-model oracle."""
    from oracle import wfm_oracle as o
    from tests.util import P
    N, Nz, B = 32, 3, 3
    out = str(tmp_path / "b0.npz")
    mp.spawn(_batch_worker, args=(2, 31500 + os.getpid() % 2000, N, Nz, B, out), nprocs=2, join=True)
    got = np.load(out)
    assert list(got["shares"]) == [0, 2]                      # rank 0 holds models 0, 1; rank 1 holds model 2
    for b in range(B):
        r = o.WideFieldModelOracle((N, N, Nz), 10, 2, P["NA"], P["lam"], P["ni"], P["dxy"], P["dz"])
        r.setPhase(got["alpha"][b]); r.setModulus(got["beta"][b])
        q = o.synthetic_q(N, N, Nz, seed=42 + b)
        assert o.rel_l2(got["p"][b], r.apply_J_phase(q)) <= 1e-12
        assert o.rel_l2(got["d"][b], r.apply_J_defocus(q)) <= 1e-12
        assert o.rel_l2(got["mo"][b], r.apply_J_modulus(q)) <= 1e-12
