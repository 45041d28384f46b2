"""z-slab sharding of the widefield PSF path across the GPUs of one box.

One process per GPU.  Planes are independent given (rho, phi, psi, Z) (WFM:291-333, 888-945), so
each rank owns the contiguous slab ``[z0, z0+nz_local)`` of the global stack and no data-path
collective is needed for the PSF.  The only exchange is one sum-allreduce of the
``3 + nPhase + nModulus`` gradient doubles per Jacobian evaluation (all three vectors ride one
message); the PSF all-gather is optional and off the hot path.  ``torch.distributed`` is plumbing:
NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

from . import _capi as capi
from .wide_field_model import WideFieldModel


def slab_bounds(nz: int, world: int, rank: int):
    """Contiguous z-slab of ``rank``: the first ``nz % world`` ranks hold one extra plane."""
    base, rem = divmod(nz, world)
    z0 = rank * base + min(rank, rem)
    return z0, base + (1 if rank < rem else 0)


def connect_peer_exchange(model, dist, group=None) -> bool:
    """Collective: map every rank's landing buffer into every rank (CUDA IPC) so that k_jac_final sums the gradient
    over the ranks by itself.  Returns True when EVERY rank succeeded; otherwise every rank is left unconnected and
    the caller falls back to the all-reduce (e.g. IPC unavailable in the container, ranks on different boxes)."""
    import torch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ok = 1
    try:
        mine = model.exchangeExport(world)
    except Exception:                                                    # noqa: BLE001
        ok, mine = 0, b""
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    if ok and all(isinstance(h, (bytes, bytearray)) and len(h) == capi.WFM_EXCHANGE_HANDLE_BYTES for h in handles):
        try:
            model.exchangeConnect(rank, world, handles)
        except Exception:                                                # noqa: BLE001
            ok = 0
    else:
        ok = 0
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    if int(t.item()) == 0:
        try:
            model.exchangeClose()
        except Exception:                                                # noqa: BLE001
            pass
        return False
    dist.barrier(group=group)               # every rank has mapped every buffer before the first store
    return True


class ShardedWideFieldModel:
    """Same call surface as WideFieldModel for the hot path; every rank passes its own slab of q."""

    def __init__(self, psfShape, nPhase, nModulus, NA, lambda_, ni, dxy, dz, radial=False, single=False, *,
                 group=None, device=0, lib=None, basis=None, exchange="nccl"):
        """``exchange``: "nccl" = one torch.distributed all-reduce of the K-vector per evaluation (any backend);
        "peer" = the sum is taken inside k_jac_final over CUDA-IPC peer memory (NVLink; wfm_exchange_*), no collective
        call at all -- needs one process per GPU on one box; "auto" = peer when every rank can set it up, else nccl."""
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.exchange = exchange
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        nz = psfShape[2] if not hasattr(psfShape, "dimension") else psfShape.dimension(2)
        self.z0, self.nz_local = slab_bounds(nz, self.world, self.rank)
        if self.nz_local <= 0:
            raise ValueError("more ranks than z-planes")
        self.model = WideFieldModel(psfShape, nPhase, nModulus, NA, lambda_, ni, dxy, dz, radial, single,
                                    device=device, z0=self.z0, nz_local=self.nz_local, lib=lib, basis=basis)
        self._handle_stream_ptr = None      # None: the handle's own (non-blocking) stream
        if exchange not in ("nccl", "peer", "auto"):
            raise ValueError("exchange must be 'nccl', 'peer' or 'auto'")
        if exchange in ("peer", "auto") and self.world > 1:
            if connect_peer_exchange(self.model, dist, group):
                self.exchange = "peer"
            elif exchange == "peer":
                raise RuntimeError("peer-memory gradient exchange could not be set up on every rank")
            else:
                self.exchange = "nccl"
        elif self.world == 1:
            self.exchange = "nccl"

    def close(self):
        if self.exchange == "peer" and self.world > 1:
            self.model.synchronize()
            self._dist.barrier(group=self.group)   # nobody unmaps while a peer may still store
            self.model.exchangeClose()
        self.model.close()

    def __getattr__(self, name):           # setters / getters are slab-local and identical on every rank
        return getattr(self.model, name)

    # ---- gradient allreduce -----------------------------------------------------------------------
    def _allreduce_host(self, vec: np.ndarray) -> np.ndarray:
        if self.world == 1 or self.exchange == "peer":       # "peer": wfm_apply_j_* already returned the global sum
            return vec
        import torch
        t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64).copy())
        backend = self._dist.get_backend(self.group)
        if backend == "nccl":
            t = t.cuda()
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def apply_J_phase(self, q_local):
        return self._allreduce_host(self.model.apply_J_phase(q_local).data)

    def apply_J_defocus(self, q_local):
        return self._allreduce_host(self.model.apply_J_defocus(q_local).data)

    def apply_J_modulus(self, q_local):
        return self._allreduce_host(self.model.apply_J_modulus(q_local).data)

    def apply_J_all(self, q_local):
        d, p, m = self.model.apply_J_all(q_local)
        full = self._allreduce_host(np.concatenate([d, p, m]))      # one message for all three
        return full[:3], full[3:3 + p.size], full[3 + p.size:]

    def applyJacobianDeviceAllReduce(self, kinds, q_tensor, grad_tensor):
        """Device-resident path: q / grad are torch CUDA tensors; NCCL allreduce of the K-vector on the current
        torch stream.  The Jacobian kernels run on the HANDLE's stream; unless that is the current torch stream
        (WideFieldModel.setStream) the collective is ordered behind them here: an event recorded on the handle's
        stream after k_jac_final, waited for by the torch stream, so NCCL can never sum a gradient that has not
        been written yet."""
        import torch
        if self.exchange == "peer":                       # the sum over the ranks happens inside k_jac_final
            self.model.applyJacobianDevice(kinds, q_tensor.data_ptr(), grad_tensor.data_ptr())
            return grad_tensor
        cur = torch.cuda.current_stream()
        if self.world > 1 and self._handle_stream_ptr != cur.cuda_stream:
            # q may have been produced on the torch stream: the handle's stream waits for it first
            self.model.waitForStream(cur.cuda_stream)
        self.model.applyJacobianDevice(kinds, q_tensor.data_ptr(), grad_tensor.data_ptr())
        if self.world > 1:
            if self._handle_stream_ptr != cur.cuda_stream:
                self.model.orderStreamAfter(cur.cuda_stream)
            self._dist.all_reduce(grad_tensor, op=self._dist.ReduceOp.SUM, group=self.group)
        return grad_tensor

    def setStream(self, cuda_stream_ptr):
        self._handle_stream_ptr = int(cuda_stream_ptr or 0) or None
        self.model.setStream(cuda_stream_ptr)

    def gatherPsfDevice(self):
        """NCCL all-gather of the PSF slabs on the device: returns a CUDA tensor (Nz, Ny, Nx) holding the
        whole stack on every rank.  Bandwidth-heavy (8*Npix*Nz bytes), therefore off the hot path."""
        import torch
        local = self.model.devicePsfTensor()
        self.model.synchronize()
        if self.world == 1:
            return local
        nz = self.model.Nz
        out = torch.empty((nz, self.model.Ny, self.model.Nx), dtype=local.dtype, device=local.device)
        bounds = [slab_bounds(nz, self.world, r) for r in range(self.world)]
        if len({b[1] for b in bounds}) == 1:
            self._dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        else:
            parts = [out[z0:z0 + n] for z0, n in bounds]
            self._dist.all_gather(parts, local.contiguous(), group=self.group)
        return out

    # ---- optional PSF gather ------------------------------------------------------------------------
    def getPsf(self, gather=False):
        local = self.model.getPsf()
        if not gather or self.world == 1:
            return local
        import torch
        parts = [None] * self.world
        self._dist.all_gather_object(parts, local, group=self.group)
        return np.concatenate(parts, axis=0)


class ShardedWideFieldModelBatch:
    """BASELINE config 5 across the GPUs of one box: ``nbatch`` INDEPENDENT models are split by model index
    (``slab_bounds(nbatch, world, rank)``), every rank holds its share on one batch handle
    (``WideFieldModelBatch``).  The models share nothing, so the data path has NO collective;
    the only exchange offered is the optional all-gather of the small per-model gradient rows."""

    def __init__(self, psfShape, nbatch, nPhase, nModulus, NA, lambda_, ni, dxy, dz, radial=False, single=False, *,
                 group=None, device=0, lib=None, basis=None):
        import torch.distributed as dist
        from .wide_field_model import WideFieldModelBatch
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nbatch = int(nbatch)
        self.b0, self.nb_local = slab_bounds(self.nbatch, self.world, self.rank)
        if self.nb_local <= 0:
            raise ValueError("more ranks than models")
        # (a batch handle needs nbatch >= 2 to differ from a plain one; one model is simply a batch of one)
        self.model = WideFieldModelBatch(psfShape, self.nb_local, nPhase, nModulus, NA, lambda_, ni, dxy, dz, radial,
                                         single, device=device, lib=lib, basis=basis)

    def __getattr__(self, name):
        return getattr(self.model, name)

    def local_rows(self, table):
        """Rows [b0, b0 + nb_local) of a global per-model table."""
        t = np.asarray(table)
        if t.shape[0] != self.nbatch:
            raise ValueError("table must have one row per model of the GLOBAL batch")
        return t[self.b0:self.b0 + self.nb_local]

    def setPhaseBatch(self, alpha_global):
        self.model.setPhaseBatch(self.local_rows(alpha_global))

    def setModulusBatch(self, beta_global):
        self.model.setModulusBatch(self.local_rows(beta_global))

    def setDefocusBatch(self, defoc_global):
        self.model.setDefocusBatch(self.local_rows(defoc_global))

    def applyJacobianBatch(self, q_local, kinds=capi.WFM_J_DEFOCUS | capi.WFM_J_PHASE | capi.WFM_J_MODULUS, gather=False):
        """q_local: this rank's models only, (nb_local, Nz, Ny, Nx).  gather=True returns the rows of ALL models."""
        d, p, m = self.model.applyJacobianBatch(q_local, kinds)
        if not gather or self.world == 1:
            return d, p, m
        parts = [None] * self.world
        self._dist.all_gather_object(parts, np.concatenate([d, p, m], axis=1), group=self.group)
        full = np.concatenate(parts, axis=0)
        return full[:, :3], full[:, 3:3 + p.shape[1]], full[:, 3 + p.shape[1]:]
