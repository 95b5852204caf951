"""
viprs_b200 -- B200-native (sm_100a) coordinate-ascent E-step for VIPRS / VIPRSMix / VIPRSGrid.

Only the hot path of shz9/viprs is here: the per-iteration CAVI sweep
(/root/reference/viprs/model/vi/e_step.hpp) plus the M-step / ELBO reductions around it, behind
the reference's own e_step boundary.  Host code holds device buffers (torch tensors) and calls
hand-written CUDA kernels through a C ABI (include/viprs_b200.h) with ctypes.  No CPU fallback.
"""
from ._lib import ViprsB200Error, lib, LIB_PATH  # noqa: F401
from .ld import DeviceLD  # noqa: F401
from .e_step import (cpp_e_step, cpp_e_step_mixture, cpp_e_step_grid, e_step_device,  # noqa: F401
                     e_step_mixture_device, e_step_grid_device, q_offset_device,
                     cpp_e_step_resident, cpp_e_step_mixture_resident,
                     e_step_incremental_device, e_step_mixture_incremental_device,
                     check_omp_support, check_blas_support)

__version__ = "0.1.0"
