"""microtipi_b200 -- B200-native (sm_100a) widefield PSF model + Jacobians of microTiPi.

Only the path BASELINE.json names lives here: the C-ABI library ``csrc/libwfm_b200.so``
(hand-written CUDA kernels, include/wfm_b200.h) and the host-side mirror of the reference's
``WideFieldModel`` / ``MicroscopeModel`` API that calls it.  Importing the package does not load
the library; constructing a model does, and fails loudly when it is missing (no CPU fallback)."""
from ._capi import load_library, LIB_PATH  # noqa: F401
from .wide_field_model import (DoubleShapedVector, DoubleShapedVectorSpace, MicroscopeModel, Shape,  # noqa: F401
                               WideFieldModel, WideFieldModelBatch)

from .convolution_cost import WeightedConvolutionCost  # noqa: F401,E402

__version__ = "0.1"
