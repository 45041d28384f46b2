"""Host-side mirror of the reference's Java API for the widefield PSF path, over the C ABI.

The reference is Java and this image has no JDK, so the class a microTiPi user sees is restated
here with the same names, argument meaning and error behaviour (IllegalArgumentException ->
ValueError), delegating every computation to ``libwfm_b200.so`` (include/wfm_b200.h):

    MicroscopeModel  <-  /root/reference/src/microTiPi/microscopy/MicroscopeModel.java:40-107
    WideFieldModel   <-  /root/reference/src/microTiPi/epifluorescence/WideFieldModel.java (WFM)

plus the three TiPi value types the callers touch (``Shape``, ``DoubleShapedVectorSpace``,
``DoubleShapedVector``; TiPi source is not in the reference tree, only the members used at
PSF_Estimation.java:117,144,202-217 are provided).  The Java binding a maintainer would write over
the same ABI is in INTEGRATION.md.  No oracle import, no numpy compute path: numpy is used only to
hold host buffers.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _capi as capi


# ---- TiPi value types (minimal) -------------------------------------------------------------------
class Shape:
    def __init__(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        self.dims = tuple(int(d) for d in dims)

    def rank(self):
        return len(self.dims)

    def dimension(self, k):
        return self.dims[k]

    def number(self):
        return int(np.prod(self.dims))


class DoubleShapedVectorSpace:
    """mitiv.linalg.shaped.DoubleShapedVectorSpace: identity of the space object is what
    apply_Jacobian / setParam dispatch on (WFM:399-422)."""

    def __init__(self, *dims):
        self.shape = Shape(*dims)

    def getNumber(self):
        return self.shape.number()

    def getShape(self):
        return self.shape

    def create(self, value=None):
        if value is None:
            data = np.zeros(self.getNumber())
        elif np.isscalar(value):
            data = np.full(self.getNumber(), float(value))
        else:
            data = np.array(value, dtype=np.float64).ravel().copy()
        return DoubleShapedVector(self, data)

    def wrap(self, arr):
        data = np.asarray(arr, dtype=np.float64).ravel()
        if data.size != self.getNumber():
            raise ValueError("array does not fit the vector space")
        return DoubleShapedVector(self, data)


class DoubleShapedVector:
    def __init__(self, space, data):
        self.space = space
        self.data = data

    def getOwner(self):
        return self.space

    def getSpace(self):
        return self.space

    def belongsTo(self, space):
        return space is self.space

    def getNumber(self):
        return self.data.size

    def get(self, i):
        return float(self.data[i])

    def set(self, i, v):
        self.data[i] = v

    def getData(self):
        return self.data

    def norm2(self):
        return math.sqrt(float(np.dot(self.data, self.data)))

    def clone(self):
        return DoubleShapedVector(self.space, self.data.copy())


def _as_host(q, dtype):
    """Host gradient array (ShapedVector or ndarray) -> contiguous buffer in reference flat order."""
    if isinstance(q, DoubleShapedVector):
        q = q.data
    return np.ascontiguousarray(q, dtype=dtype)


# ---- MicroscopeModel ----------------------------------------------------------------------------
class MicroscopeModel:
    """MicroscopeModel.java:40-107."""

    NORMALIZED = True
    DEUXPI = 2 * math.pi

    def __init__(self, psfShape, dxy, dz, single):
        if not isinstance(psfShape, Shape):
            psfShape = Shape(psfShape)
        if psfShape.rank() != 3:
            raise ValueError("Microscope PSF  should be 3D")               # MicroscopeModel.java:70-72
        self.PState = 0
        self.dxy, self.dz = float(dxy), float(dz)
        self.Nx, self.Ny, self.Nz = psfShape.dimension(0), psfShape.dimension(1), psfShape.dimension(2)
        self.psfShape = psfShape
        self.single = bool(single)
        self.psf = None
        self.parameterSpace = None
        self.parameterCoefs = None

    def isSingle(self):
        return self.single

    def setSingle(self, single):
        self.single = bool(single)

    def getShape(self):
        return self.psfShape

    def apply_Jacobian(self, grad, xspace):      # abstract, MicroscopeModel.java:90
        raise NotImplementedError

    def getParametersFlags(self):                # abstract, MicroscopeModel.java:96
        raise NotImplementedError

    def computePsf(self):                        # abstract, MicroscopeModel.java:103
        raise NotImplementedError


# ---- WideFieldModel -----------------------------------------------------------------------------
class WideFieldModel(MicroscopeModel):
    DEFOCUS, PHASE, MODULUS = 0, 1, 2            # WFM:113-121
    parametersFlag = [0, 1, 2]                   # WFM:123

    def __init__(self, psfShape, nPhase=0, nModulus=1, NA=None, lambda_=None, ni=None, dxy=None, dz=None,
                 radial=False, single=False, *, device=0, z0=0, nz_local=None, lib=None, basis=None, nbatch=1,
                 devices=None):
        """WFM:154-188.  ``z0/nz_local`` make this object one z-slab of the global stack;
        ``lib`` lets the tests bind another build of the same ABI; ``basis`` (optional
        ``callable(Nzern) -> Z[Nzern, Npix]``) replaces the device-side computeZernike(); ``nbatch`` > 1 makes the
        handle a batch of independent models (see WideFieldModelBatch); ``devices`` (a list of CUDA device indices)
        spreads the stack over several GPUs of the box behind the same calls (wfm_create_multi)."""
        super().__init__(psfShape, dxy, dz, single)
        self._lib = lib if lib is not None else capi.load_library()
        self._h = C.c_void_p()
        self._basis_fn = basis
        nzl = self.Nz - z0 if nz_local is None else nz_local
        self.z0, self.nz_local = int(z0), int(nzl)
        self.nbatch = int(nbatch)
        prec = capi.WFM_F32 if single else capi.WFM_F64
        self.devices = None if devices is None else [int(d) for d in devices]
        if self.devices is not None:
            if self.nbatch > 1 or z0 != 0 or (nz_local is not None and nz_local != self.Nz):
                raise ValueError("a multi-device model holds the whole stack of one model")
            arr = (C.c_int * len(self.devices))(*self.devices)
            rc = self._lib.wfm_create_multi(C.byref(self._h), self.Nx, self.Ny, self.Nz, self.dxy, self.dz, prec, arr,
                                            len(self.devices))
        elif self.nbatch > 1:
            rc = self._lib.wfm_create_batch(C.byref(self._h), self.Nx, self.Ny, self.Nz, self.nbatch, self.dxy,
                                            self.dz, prec, int(device))
            self.nz_local = self.Nz * self.nbatch                          # planes held by the handle
        else:
            rc = self._lib.wfm_create_slab(C.byref(self._h), self.Nx, self.Ny, self.Nz, self.z0, self.nz_local,
                                           self.dxy, self.dz, prec, int(device))
        if rc != capi.WFM_OK:
            msg = self._lib.wfm_last_error(None).decode()
            self._h = C.c_void_p()
            if rc in (capi.WFM_ERR_INVALID_ARG, capi.WFM_ERR_UNSUPPORTED):
                raise ValueError(msg)                                      # WFM:158-160
            raise RuntimeError(msg)
        self.lambda_ = float(lambda_)
        self.ni = float(ni)
        self.Nzern = 4                                                     # WFM:163
        self.NA = float(NA)
        self.radius = self.NA / self.lambda_                               # WFM:165
        self.lambda_ni = self.ni / self.lambda_                            # WFM:166
        self.deltaX = self.deltaY = 0.0
        self.radial = bool(radial)
        self._call("wfm_set_optics", self.NA, self.lambda_, self.ni)       # computeMaskPupil() WFM:174
        self.nModulus = max(1, int(nModulus))                              # WFM:176-179
        self.parameterSpace = [None, None, None]                           # WFM:181-182
        self.parameterCoefs = [None, None, None]
        self.nPhase = int(nPhase)
        self._setNModulus()                                                # WFM:185
        self._setNPhase()                                                  # WFM:186
        self._setDefocusInner()                                            # WFM:187

    # -- plumbing --------------------------------------------------------------------------------
    def _call(self, name, *args):
        rc = getattr(self._lib, name)(self._h, *args)
        if rc != capi.WFM_OK:
            msg = self._lib.wfm_last_error(self._h).decode()
            if rc == capi.WFM_ERR_INVALID_ARG:
                raise ValueError(msg)
            raise RuntimeError(f"{name}: {msg} (status {rc})")
        return rc

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.wfm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dtype(self):
        return np.float32 if self.single else np.float64

    def _npix(self):
        return self.Nx * self.Ny

    def setStream(self, cuda_stream_ptr):
        self._call("wfm_set_stream", C.c_void_p(cuda_stream_ptr))

    def synchronize(self):
        self._call("wfm_synchronize")

    def waitForStream(self, cuda_stream_ptr):
        """The handle's stream waits (on the device) for the work queued on another stream so far."""
        self._call("wfm_wait_stream", C.c_void_p(cuda_stream_ptr))

    def orderStreamAfter(self, cuda_stream_ptr):
        """Another stream waits (on the device) for the work queued on the handle's stream so far."""
        self._call("wfm_fence_stream", C.c_void_p(cuda_stream_ptr))

    # -- basis -----------------------------------------------------------------------------------
    def computeZernike(self):                                              # WFM:194-197
        if self._basis_fn is not None:
            Z = np.ascontiguousarray(self._basis_fn(self.Nzern), dtype=np.float64)
            self._call("wfm_set_basis", Z.ctypes.data_as(C.c_void_p), self.Nzern, int(self.radial))
        else:
            self._call("wfm_build_basis", self.Nzern, int(self.radial))

    # -- hot path ----------------------------------------------------------------------------------
    def computePsf(self):                                                  # WFM:206-396
        if self.PState > 0:
            return
        self._call("wfm_compute_psf")
        self.PState = 1

    def apply_Jacobian(self, grad, xspace):                                # WFM:399-409
        if xspace is self.parameterSpace[self.DEFOCUS] and xspace is not None:
            return self.apply_J_defocus(grad)
        if xspace is self.parameterSpace[self.PHASE] and xspace is not None:
            return self.apply_J_phase(grad)
        if xspace is self.parameterSpace[self.MODULUS] and xspace is not None:
            return self.apply_J_modulus(grad)
        raise ValueError("DoubleShapedVector grad does not belong to any space")

    def setParam(self, param):                                             # WFM:412-422, 1553-1556
        if not isinstance(param, DoubleShapedVector):
            return self.setDefocus(param)
        if param.getOwner() is self.parameterSpace[self.DEFOCUS]:
            self.setDefocus(param)
        elif param.getOwner() is self.parameterSpace[self.PHASE]:
            self.setPhase(param)
        elif param.getOwner() is self.parameterSpace[self.MODULUS]:
            self.setModulus(param)
        else:
            raise ValueError("DoubleShapedVector param does not belong to any space")

    def _apply(self, name, q, space):
        qh = _as_host(q, self._dtype())
        if qh.size != self._npix() * self.nz_local:
            raise ValueError("gradient does not have the shape of the PSF")
        out = np.zeros(space.getNumber())
        self._call(name, qh.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), out.size)
        self.PState = self._lib.wfm_psf_state(self._h)
        return space.create(out)

    def apply_J_modulus(self, q):                                          # WFM:429-730
        return self._apply("wfm_apply_j_modulus", q, self.parameterSpace[self.MODULUS])

    def apply_J_phase(self, q):                                            # WFM:738-1021
        if self.parameterSpace[self.PHASE] is None:
            raise ValueError("phase space is empty")
        return self._apply("wfm_apply_j_phase", q, self.parameterSpace[self.PHASE])

    def apply_J_defocus(self, q):                                          # WFM:1029-1369
        return self._apply("wfm_apply_j_defocus", q, self.parameterSpace[self.DEFOCUS])

    def apply_J_all(self, q):
        """All three Jacobians from one adjoint FFT pass (they differ only after the FFT)."""
        qh = _as_host(q, self._dtype())
        if qh.size != self._npix() * self.nz_local:
            raise ValueError("gradient does not have the shape of the PSF")
        d = np.zeros(3)
        p = np.zeros(max(self.getNPhase(), 1))
        m = np.zeros(self.getNModulus())
        self._call("wfm_apply_j_all", qh.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                   p.ctypes.data_as(C.c_void_p) if self.getNPhase() else None, m.ctypes.data_as(C.c_void_p))
        self.PState = self._lib.wfm_psf_state(self._h)
        return d, p[:self.getNPhase()], m

    # -- device-resident entry points (benchmark / multi-GPU plumbing) ---------------------------------
    def gradLength(self):
        return int(self._lib.wfm_grad_length(self._h))

    def applyJacobianDevice(self, kinds, q_dev_ptr, grad_dev_ptr):
        """wfm_apply_jacobian_dev: q and grad are raw device pointers; asynchronous on the stream.
        grad layout: [defocus(3) | phase(nPhase) | modulus(nModulus)], this slab's partial sums."""
        self._call("wfm_apply_jacobian_dev", int(kinds), C.c_void_p(q_dev_ptr), C.c_void_p(grad_dev_ptr))
        self.PState = self._lib.wfm_psf_state(self._h)

    def devicePsfPointer(self):
        p = C.c_void_p()
        self._call("wfm_device_psf", C.byref(p))
        self.PState = 1
        return p.value

    def deviceCpxPsfPointer(self):
        p = C.c_void_p()
        self._call("wfm_device_cpx_psf", C.byref(p))
        self.PState = 1
        return p.value

    def devicePsfTensor(self):
        """Zero-copy torch view (nz_local, Ny, Nx) of the device-resident PSF slab (plumbing for NCCL)."""
        import torch

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (self.nz_local, self.Ny, self.Nx), "typestr": "<f4" if self.single else "<f8",
                                      "data": (self.devicePsfPointer(), False), "version": 2}
        return torch.as_tensor(v, device="cuda")

    # -- multi-device handles (wfm_create_multi) ------------------------------------------------------------------
    def parts(self):
        """[(device, z0, nz_local, child handle)] of a multi-device model ([] for a plain one)."""
        out = []
        for i in range(self._lib.wfm_multi_parts(self._h)):
            d, z, n = C.c_int(), C.c_int(), C.c_int()
            child = C.c_void_p()
            self._call("wfm_multi_part_info", i, C.byref(d), C.byref(z), C.byref(n))
            self._call("wfm_multi_part", i, C.byref(child))
            out.append((d.value, z.value, n.value, child))
        return out

    def applyJacobianDeviceMulti(self, kinds, q_dev_ptrs, grad_dev_ptr):
        """wfm_multi_apply_jacobian_dev: one device pointer per part (its slab of q, on its device); the summed
        gradient lands on the first device.  Asynchronous."""
        arr = (C.c_void_p * len(q_dev_ptrs))(*[C.c_void_p(p) for p in q_dev_ptrs])
        self._call("wfm_multi_apply_jacobian_dev", int(kinds), arr, C.c_void_p(grad_dev_ptr))
        self.PState = self._lib.wfm_psf_state(self._h)

    # -- cross-process gradient exchange over peer memory (one process per GPU) -------------------------------------
    def exchangeExport(self, world):
        buf = (C.c_char * capi.WFM_EXCHANGE_HANDLE_BYTES)()
        self._call("wfm_exchange_export", int(world), buf)
        return bytes(buf)

    def exchangeConnect(self, rank, world, handles):
        blob = b"".join(handles)
        if len(blob) != world * capi.WFM_EXCHANGE_HANDLE_BYTES:
            raise ValueError("one 64-byte handle per rank")
        self._call("wfm_exchange_connect", int(rank), int(world), C.c_char_p(blob))

    def exchangeStatus(self):
        rc = self._lib.wfm_exchange_status(self._h)
        if rc < 0:
            raise RuntimeError(self._lib.wfm_last_error(self._h).decode())
        return rc

    def exchangeClose(self):
        self._call("wfm_exchange_close")

    def fillUniform(self, dev_ptr, seed, first_index, count, single=None):
        prec = capi.WFM_F32 if (self.single if single is None else single) else capi.WFM_F64
        self._call("wfm_fill_uniform", C.c_void_p(dev_ptr), prec, int(seed), int(first_index), int(count))

    def setProfiling(self, on):
        self._call("wfm_set_profiling", 1 if on else 0)

    def kernelTimes(self):
        """{kernel group: (total ms, launch groups)} since profiling was switched on."""
        n = len(capi.KERNEL_NAMES)
        ms = (C.c_double * n)()
        cnt = (C.c_uint64 * n)()
        self._call("wfm_get_kernel_times", ms, cnt, n)
        return {capi.KERNEL_NAMES[i]: (ms[i], int(cnt[i])) for i in range(n)}

    def activeExtent(self):
        ax, ay = C.c_int(), C.c_int()
        self._call("wfm_active_extent", C.byref(ax), C.byref(ay))
        return ax.value, ay.value

    # -- defocus -----------------------------------------------------------------------------------
    def computeDefocus(self):                                              # WFM:1452-1499
        self._call("wfm_set_defocus", (C.c_double * 3)(self.lambda_ni, self.deltaX, self.deltaY), 3)

    def setDefocus(self, defoc):                                           # WFM:1510-1534 / 1543-1549
        if not isinstance(defoc, DoubleShapedVector):
            if self.parameterSpace[self.DEFOCUS] is None:
                self.parameterSpace[self.DEFOCUS] = DoubleShapedVectorSpace(3)
            defoc = self.parameterSpace[self.DEFOCUS].wrap(defoc)
        if not defoc.belongsTo(self.parameterSpace[self.DEFOCUS]):
            raise ValueError("defocus  does not belong to the parameterSpace[DEFOCUS]")
        n = defoc.getNumber()
        if n not in (1, 3):
            raise ValueError("bad defocus  parameters")                    # WFM:1530 (+ quirk Q4 for n == 2)
        self.parameterCoefs[self.DEFOCUS] = defoc
        if n == 3:
            self.deltaX, self.deltaY = defoc.get(1), defoc.get(2)
        self.lambda_ni = defoc.get(0)
        self.ni = self.lambda_ni * self.lambda_
        arr = (C.c_double * n)(*[defoc.get(i) for i in range(n)])
        self._call("wfm_set_defocus", arr, n)
        self.freeMem()

    def _setDefocusInner(self):                                            # WFM:1562-1564
        self.setDefocus([self.ni / self.lambda_, self.deltaX, self.deltaY])

    def setPupilAxis(self, axis):                                          # WFM:1573-1579
        self.setDefocus([self.ni / self.lambda_, axis[0], axis[1]])

    def setNi(self, value):                                                # WFM:1698-1707
        self.ni = float(value)
        self.lambda_ni = self.ni / self.lambda_
        self.setDefocus([self.ni / self.lambda_, self.deltaX, self.deltaY])

    # -- modulus / phase ---------------------------------------------------------------------------
    def setModulus(self, modulus):                                         # WFM:1588-1610 / 1616-1620
        if not isinstance(modulus, DoubleShapedVector):
            modulus = np.asarray(modulus, dtype=np.float64)
            self.setNModulus(modulus.size)
            modulus = self.parameterSpace[self.MODULUS].wrap(modulus)
        if not modulus.belongsTo(self.parameterSpace[self.MODULUS]):
            raise ValueError("DoubleShapedVector beta does not belong to the modulus space")
        self.parameterCoefs[self.MODULUS] = modulus
        b = np.ascontiguousarray(modulus.data, dtype=np.float64)
        self._call("wfm_set_modulus", b.ctypes.data_as(C.c_void_p), b.size)
        self.freeMem()

    def setPhase(self, phase):                                             # WFM:1625-1649 / 1655-1665
        if not isinstance(phase, DoubleShapedVector):
            if phase is None or len(phase) == 0:
                self.nPhase = 0
                self.parameterSpace[self.PHASE] = None
                self.parameterCoefs[self.PHASE] = None
                self._call("wfm_set_phase", None, 0)                       # the device side drops its phase vector too
                self.freeMem()
                return
            phase = np.asarray(phase, dtype=np.float64)
            self.setNPhase(phase.size)
            phase = self.parameterSpace[self.PHASE].wrap(phase)
        if self.parameterSpace[self.PHASE] is None or not phase.belongsTo(self.parameterSpace[self.PHASE]):
            raise ValueError("phase parameter does not belong to the right space  ")
        self.parameterCoefs[self.PHASE] = phase
        a = np.ascontiguousarray(phase.data, dtype=np.float64)
        self._call("wfm_set_phase", a.ctypes.data_as(C.c_void_p), a.size)
        self.freeMem()

    def _setNPhase(self):                                                  # WFM:1899-1914
        if self.nPhase > 0:
            self.parameterSpace[self.PHASE] = DoubleShapedVectorSpace(self.nPhase)
            off = 1 if self.radial else 3
            self.Nzern = max(self.nPhase + off, self.parameterSpace[self.MODULUS].getNumber())
            self.computeZernike()
            self.parameterCoefs[self.PHASE] = self.parameterSpace[self.PHASE].create(0.0)
            self.setPhase(self.parameterCoefs[self.PHASE])
        else:
            self.parameterSpace[self.PHASE] = None
            self.parameterCoefs[self.PHASE] = None

    def setNPhase(self, nPh):                                              # WFM:1919-1922
        self.nPhase = int(nPh)
        self._setNPhase()

    def setNModulus(self, nMod):                                           # WFM:1930-1934
        self.nModulus = int(nMod)
        self._setNModulus()

    def _setNModulus(self):                                                # WFM:1939-1961
        if self.nModulus < 1:
            self.nModulus = 1
        self.parameterSpace[self.MODULUS] = DoubleShapedVectorSpace(self.nModulus)
        if self.parameterSpace[self.PHASE] is None:
            self.Nzern = self.nModulus
        else:
            off = 1 if self.radial else 3
            self.Nzern = max(self.parameterSpace[self.PHASE].getNumber() + off, self.nModulus)
        self.computeZernike()
        self.parameterCoefs[self.MODULUS] = self.parameterSpace[self.MODULUS].create(0.0)
        self.parameterCoefs[self.MODULUS].set(0, 1.0)
        self.setModulus(self.parameterCoefs[self.MODULUS])

    def setModulusMode(self, last_plane_only: bool):
        """Quirk Q1 switch: False = intended (sum over z), True = live fp64 reference behaviour."""
        self._call("wfm_set_modulus_mode", capi.WFM_MODULUS_REFERENCE_LAST_PLANE if last_plane_only
                   else capi.WFM_MODULUS_INTENDED)

    def setPupilArrays(self, rho=None, phi=None, psi=None, mask=None):
        """Escape hatch of the ABI: load identical synthetic pupils verbatim."""
        def p(a, dt):
            if a is None:
                return None, None
            arr = np.ascontiguousarray(a, dtype=dt)
            return arr, arr.ctypes.data_as(C.c_void_p)
        keep = [p(rho, np.float64), p(phi, np.float64), p(psi, np.float64), p(mask, np.uint8)]
        self._call("wfm_set_pupil_arrays", *[k[1] for k in keep])
        self.freeMem()

    # -- getters -----------------------------------------------------------------------------------
    def _get_pupil(self, name, dtype=np.float64):
        if self.PState < 1:
            self.computePsf()                                              # WFM:1674-1676 etc.
        out = np.empty(self._npix() * self.nbatch, dtype=dtype)
        self._call(name, out.ctypes.data_as(C.c_void_p))
        return out if self.nbatch == 1 else out.reshape(self.nbatch, -1)

    def getRho(self):                                                      # WFM:1673
        return self._get_pupil("wfm_get_rho")

    def getPhi(self):                                                      # WFM:1713
        return self._get_pupil("wfm_get_phi")

    def getPsi(self):                                                      # WFM:1723
        return self._get_pupil("wfm_get_psi")

    def getMaskPupil(self):                                                # WFM:1784
        return self._get_pupil("wfm_get_mask", np.uint8).astype(bool)

    def getLambda(self):
        return self.lambda_

    def getNi(self):
        return self.ni

    def getModulusCoefs(self):
        return self.parameterCoefs[self.MODULUS]

    def getPhaseCoefs(self):
        return self.parameterCoefs[self.PHASE]

    def getDefocusMultiplyByLambda(self):                                  # WFM:1750-1756
        if self.PState < 1:
            self.computePsf()
        return [self.lambda_ni * self.lambda_, self.deltaX * self.lambda_, self.deltaY * self.lambda_]

    def getDefocus(self):                                                  # WFM:1761-1767
        if self.PState < 1:
            self.computePsf()
        return [self.lambda_ni, self.deltaX, self.deltaY]

    def getPupilShift(self):                                               # WFM:1772-1778
        if self.PState < 1:
            self.computePsf()
        return [self.deltaX, self.deltaY]

    def getPsf(self):                                                      # WFM:1798-1804
        """Host copy with numpy shape (nz_local, Ny, Nx): ``.ravel()`` is the reference flat order."""
        if self.PState < 1:
            self.computePsf()
        out = np.empty((self.nz_local, self.Ny, self.Nx), dtype=self._dtype())
        self._call("wfm_get_psf", out.ctypes.data_as(C.c_void_p))
        self.psf = out
        return out

    def getPsfAsync(self, out_ptr):
        """getPsf() into a pinned host buffer (raw address) without waiting for the copy: it runs on the
        handle's second stream beside the next host->device transfer.  Pair with waitTransfers()."""
        self._call("wfm_get_psf_async", C.c_void_p(out_ptr))
        self.PState = 1

    def waitTransfers(self):
        self._call("wfm_wait_transfers")

    def get_cpxPsf(self):                                                  # WFM:1856-1861
        if self.PState < 1:
            self.computePsf()
        out = np.empty((self.nz_local, self.Ny, self.Nx, 2), dtype=self._dtype())
        self._call("wfm_get_cpx_psf", out.ctypes.data_as(C.c_void_p))
        return out

    def getMtf(self):                                                      # WFM:1807-1828
        """The 3-D DFT of the PSF, shape (Nz, Ny, Nx, 2) -- what the reference's getMtf() is written to return
        (its copy loop `i = i++` never terminates, quirk Q8; the intended result is implemented)."""
        out = np.empty((self.nz_local, self.Ny, self.Nx, 2), dtype=np.float64)
        self._call("wfm_get_mtf", out.ctypes.data_as(C.c_void_p))
        self.PState = 1
        return out

    def getPsfRolled(self):
        """ArrayUtils.roll(getPsf()) (BlindDeconvJob.java:100): the PSF centred in the volume."""
        out = np.empty((self.nz_local, self.Ny, self.Nx), dtype=self._dtype())
        self._call("wfm_get_psf_rolled", out.ctypes.data_as(C.c_void_p))
        self.PState = 1
        return out

    def getZernike(self, k=None):                                          # WFM:1834 / 1849
        Z = np.empty((self.Nzern, self._npix()))
        self._call("wfm_get_basis", Z.ctypes.data_as(C.c_void_p), self.Nzern)
        return Z if k is None else Z[k].reshape(self.Ny, self.Nx)

    def getNZern(self):
        return self.Nzern

    def getNModulus(self):                                                 # WFM:1981
        return self.parameterCoefs[self.MODULUS].getNumber()

    def getNPhase(self):                                                   # WFM:1988-1993
        return 0 if self.parameterCoefs[self.PHASE] is None else self.parameterCoefs[self.PHASE].getNumber()

    def getParametersFlags(self):                                          # WFM:2000
        return self.parametersFlag

    def freeMem(self):                                                     # WFM:1970-1974
        self.PState = 0
        self.psf = None
        if self._h:
            self._lib.wfm_invalidate(self._h)


class WideFieldModelBatch(WideFieldModel):
    """``nbatch`` independent WideFieldModels of one shape on ONE handle (BASELINE config 5: parameter estimation over
    many bead PSFs).  The models share optics and basis; ``setPhaseBatch / setModulusBatch / setDefocusBatch`` take one
    row per model, the inherited scalar setters give every model the same vector.  ``getPsf()`` has shape
    (nbatch, Nz, Ny, Nx); ``applyJacobianBatch`` returns (defocus[nbatch,3], phase[nbatch,nPhase], modulus[nbatch,nModulus])
    -- per model exactly what apply_J_defocus / apply_J_phase / apply_J_modulus (WFM:1029, 738, 429) return."""

    def __init__(self, psfShape, nbatch, *args, **kw):
        super().__init__(psfShape, *args, nbatch=nbatch, **kw)

    def _table(self, tab, n=None):
        t = np.ascontiguousarray(tab, dtype=np.float64)
        if t.ndim != 2 or t.shape[0] != self.nbatch or (n is not None and t.shape[1] != n):
            raise ValueError("coefficient table must have one row per model")
        return t

    def setPhaseBatch(self, alpha):                                        # setPhase (WFM:1625-1649) per model
        t = self._table(alpha)
        if t.shape[1] != self.nPhase:
            self.setNPhase(t.shape[1])
        self._call("wfm_batch_set_phase", t.ctypes.data_as(C.c_void_p), t.shape[1])
        self.freeMem()

    def setModulusBatch(self, beta):                                       # setModulus (WFM:1588-1610) per model
        t = self._table(beta)
        if t.shape[1] != self.nModulus:
            self.setNModulus(t.shape[1])
        self._call("wfm_batch_set_modulus", t.ctypes.data_as(C.c_void_p), t.shape[1])
        self.freeMem()

    def setDefocusBatch(self, defoc):                                      # setDefocus (WFM:1510-1534) per model
        t = self._table(defoc)
        if t.shape[1] not in (1, 3):
            raise ValueError("bad defocus  parameters")
        self._call("wfm_batch_set_defocus", t.ctypes.data_as(C.c_void_p), t.shape[1])
        self.freeMem()

    def getPsf(self):
        return super().getPsf().reshape(self.nbatch, self.Nz, self.Ny, self.Nx)

    def get_cpxPsf(self):
        return super().get_cpxPsf().reshape(self.nbatch, self.Nz, self.Ny, self.Nx, 2)

    def applyJacobianBatch(self, q, kinds=capi.WFM_J_DEFOCUS | capi.WFM_J_PHASE | capi.WFM_J_MODULUS):
        qh = _as_host(q, self._dtype())
        if qh.size != self._npix() * self.nz_local:
            raise ValueError("gradient does not have the shape of the PSF batch")
        out = np.zeros((self.nbatch, self.gradLength()))
        self._call("wfm_batch_apply_jacobian", int(kinds), qh.ctypes.data_as(C.c_void_p),
                   out.ctypes.data_as(C.c_void_p))
        self.PState = self._lib.wfm_psf_state(self._h)
        nP, nM = C.c_int(), C.c_int()                                      # split by the handle's own counts
        self._call("wfm_get_info", None, None, None, None, None, None, None, C.byref(nP), C.byref(nM))
        return out[:, :3], out[:, 3:3 + nP.value], out[:, 3 + nP.value:3 + nP.value + nM.value]
