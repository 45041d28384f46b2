"""Host-side mirror of TiPi's ``mitiv.conv.WeightedConvolutionCost`` as microTiPi drives it
(PSF_Estimation.java:144-150 build / setPSF / setData / setWeights, :157,206 computeCostAndGradient),
over the C ABI (include/wfm_b200.h, ``wfm_conv_*``) -- SURVEY.md section 8, "next" row f1.

In PSF_Estimation the roles are swapped with respect to a deconvolution: the *object* is loaded as
the kernel of the operator (``fdata.setPSF(objArray, off)``) and the microscope PSF is the variable.
TiPi's source is not in the reference tree, so the semantics (cost = alpha/2 sum w (obj (*) h - y)^2,
periodic 3-D convolution at the data shape, offset {0,0,0}) are restated: parity unpinned.
No oracle import, no numpy compute path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from .wide_field_model import DoubleShapedVector, DoubleShapedVectorSpace, Shape


class WeightedConvolutionCost:
    def __init__(self, space, *, device=0, lib=None, devices=None):
        self._lib = lib if lib is not None else capi.load_library()
        self.space = space
        shape = space.getShape() if hasattr(space, "getShape") else Shape(space)
        if shape.rank() != 3:
            raise ValueError("the data space must be 3D")
        self.Nx, self.Ny, self.Nz = shape.dimension(0), shape.dimension(1), shape.dimension(2)
        self._h = C.c_void_p()
        if devices is not None:             # z-slabs over several GPUs (pairs with WideFieldModel(devices=...))
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._lib.wfm_conv_create_multi(C.byref(self._h), self.Nx, self.Ny, self.Nz, capi.WFM_F64, arr, len(devices))
        else:
            rc = self._lib.wfm_conv_create(C.byref(self._h), self.Nx, self.Ny, self.Nz, capi.WFM_F64, int(device))
        if rc != capi.WFM_OK:
            msg = self._lib.wfm_conv_last_error(None).decode()
            self._h = C.c_void_p()
            if rc in (capi.WFM_ERR_INVALID_ARG, capi.WFM_ERR_UNSUPPORTED):
                raise ValueError(msg)
            raise RuntimeError(msg)

    @classmethod
    def build(cls, space, **kw):                                           # PSF_Estimation.java:147
        return cls(space, **kw)

    def _call(self, name, *args):
        rc = getattr(self._lib, name)(self._h, *args)
        if rc != capi.WFM_OK:
            msg = self._lib.wfm_conv_last_error(self._h).decode()
            if rc == capi.WFM_ERR_INVALID_ARG:
                raise ValueError(msg)
            raise RuntimeError(f"{name}: {msg} (status {rc})")

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.wfm_conv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _vol(self, a):
        if isinstance(a, DoubleShapedVector):
            a = a.data
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size != self.Nx * self.Ny * self.Nz:
            raise ValueError("array does not have the shape of the data space")
        return a

    def setPSF(self, obj, off=(0, 0, 0)):                                  # PSF_Estimation.java:145,148
        if tuple(int(o) for o in off) != (0, 0, 0):
            raise ValueError("only the offset {0,0,0} used by PSF_Estimation is supported")
        a = self._vol(obj)
        self._call("wfm_conv_set_object", a.ctypes.data_as(C.c_void_p))

    def setData(self, data):                                               # :149
        a = self._vol(data)
        self._call("wfm_conv_set_data", a.ctypes.data_as(C.c_void_p))

    def setWeights(self, weights, copy=True):                              # :150
        if weights is None:
            self._call("wfm_conv_set_weights", None)
            return
        a = self._vol(weights)
        self._call("wfm_conv_set_weights", a.ctypes.data_as(C.c_void_p))

    def computeCostAndGradient(self, alpha, x, gx, clr):                   # :157,206
        """Returns the cost; writes (clr) or accumulates (not clr) the gradient into ``gx``."""
        xh = self._vol(x)
        g = gx.data if isinstance(gx, DoubleShapedVector) else gx
        if not (isinstance(g, np.ndarray) and g.dtype == np.float64 and g.flags.c_contiguous and g.size == xh.size):
            raise ValueError("gx must be a contiguous float64 array / DoubleShapedVector of the data shape")
        cost = C.c_double()
        self._call("wfm_conv_cost_and_gradient", float(alpha), xh.ctypes.data_as(C.c_void_p),
                   g.ctypes.data_as(C.c_void_p), 1 if clr else 0, C.byref(cost))
        return cost.value

    def evalFG(self, model, param_flag, x, alpha=1.0):
        """One COMPUTE_FG step of PSF_Estimation.fitPSF (PSF_Estimation.java:202-217) on the device:
        setParam(x) -> computePsf -> computeCostAndGradient -> apply_Jacobian.  Returns (cost, gradient)."""
        xv = np.ascontiguousarray(x.data if isinstance(x, DoubleShapedVector) else x, dtype=np.float64)
        if int(param_flag) not in (model.DEFOCUS, model.PHASE, model.MODULUS):
            raise ValueError("DoubleShapedVector param does not belong to any space")
        # the host mirror takes the step setParam(x) takes (WFM:412-422: x is stored in parameterCoefs, and for the
        # defocus group ni / lambda_ni / deltaX / deltaY follow, WFM:1516-1531); its setter call IS the first link of
        # the device chain, so wfm_eval_fg gets x = NULL ("parameters already set")
        if int(param_flag) == model.PHASE and xv.size == 0:
            raise ValueError("phase space is empty")
        setter = {model.DEFOCUS: model.setDefocus, model.PHASE: model.setPhase, model.MODULUS: model.setModulus}
        owner = model.parameterSpace[int(param_flag)]
        if isinstance(x, DoubleShapedVector) and x.getOwner() is owner:
            setter[int(param_flag)](x)
        elif owner is not None and owner.getNumber() == xv.size:
            setter[int(param_flag)](owner.wrap(xv.copy()))     # setParam(DoubleShapedVector): no basis rebuild (WFM:412-422)
        else:
            setter[int(param_flag)](xv.copy())                 # a new length: the double[] overloads resize the space first
        g = np.zeros(xv.size)
        cost = C.c_double()
        rc = self._lib.wfm_eval_fg(model.handle, self._h, int(param_flag), None, xv.size,
                                   float(alpha), C.byref(cost), g.ctypes.data_as(C.c_void_p))
        if rc != capi.WFM_OK:
            msg = self._lib.wfm_last_error(model.handle).decode()
            if rc == capi.WFM_ERR_INVALID_ARG:
                raise ValueError(msg)
            raise RuntimeError(f"wfm_eval_fg: {msg} (status {rc})")
        model.PState = 1
        return cost.value, g
