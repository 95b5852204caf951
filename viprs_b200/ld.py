"""
Device-resident LD matrix (the B200 counterpart of what ``VIPRS.__init__`` keeps in host RAM:
``ld_data / ld_indptr / ld_left_bound``, /root/reference/viprs/model/VIPRS.py:153-172).
"""
import ctypes

import numpy as np
import torch

from . import _lib

_NP_DT = {np.dtype(np.int8): _lib.I8, np.dtype(np.int16): _lib.I16,
          np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64}
_TORCH_DT = {torch.int8: _lib.I8, torch.int16: _lib.I16, torch.float32: _lib.F32, torch.float64: _lib.F64}


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceLD:
    """
    An LD matrix in magenpy's "CSR without column indices" layout, re-laid on the current CUDA
    device for the sweep kernels.  Accepts numpy arrays (host) or torch CUDA tensors.

    :param ld_data: 1-D int8 / int16 / float32 / float64 array of stored entries.
    :param ld_indptr: (M+1,) int32 / int64 row pointers.
    :param ld_left_bound: (M,) int32 first column of each row's contiguous run.
    """

    def __init__(self, ld_data, ld_indptr, ld_left_bound, stage_bytes=0):
        L = _lib.lib()
        if L.viprs_b200_device_count() <= 0:
            raise _lib.ViprsB200Error(-5, "DeviceLD")
        self._h = ctypes.c_void_p()
        on_dev = isinstance(ld_data, torch.Tensor)
        if on_dev:
            assert ld_data.is_cuda and ld_indptr.is_cuda and ld_left_bound.is_cuda
            ld_data = ld_data.contiguous()
            ld_indptr = ld_indptr.contiguous()
            ld_left_bound = ld_left_bound.to(torch.int32).contiguous()
            dt = _TORCH_DT[ld_data.dtype]
            is64 = int(ld_indptr.dtype == torch.int64)
            assert ld_indptr.dtype in (torch.int32, torch.int64)
            M = ld_left_bound.numel()
            args = (ld_left_bound.data_ptr(), ld_indptr.data_ptr(), is64, ld_data.data_ptr(), dt, _lib.MEM_DEVICE)
            keep = (ld_data, ld_indptr, ld_left_bound)
        else:
            ld_data = np.ascontiguousarray(ld_data)
            ld_indptr = np.ascontiguousarray(ld_indptr)
            ld_left_bound = np.ascontiguousarray(ld_left_bound, dtype=np.int32)
            dt = _NP_DT[ld_data.dtype]
            assert ld_indptr.dtype in (np.int32, np.int64)
            is64 = int(ld_indptr.dtype == np.int64)
            M = ld_left_bound.shape[0]
            args = (ld_left_bound.ctypes.data, ld_indptr.ctypes.data, is64, ld_data.ctypes.data, dt, _lib.MEM_HOST)
            keep = (ld_data, ld_indptr, ld_left_bound)
        assert ld_indptr.shape[0] == M + 1
        rc = L.viprs_b200_ld_create(ctypes.byref(self._h), M, args[0], args[1], args[2], args[3], args[4],
                                    args[5], int(stage_bytes), _stream_ptr())
        del keep
        _lib.check(rc, "viprs_b200_ld_create")
        info = _lib.LdInfo()
        _lib.check(L.viprs_b200_ld_info(self._h, ctypes.byref(info)), "viprs_b200_ld_info")
        self.M = info.M
        self.ld_dtype = info.ld_dtype
        self.n_blocks = info.n_blocks
        self.max_block = info.max_block
        self.n_panels = info.n_panels
        self.stage_bytes = info.stage_bytes
        self.nnz = info.nnz
        self.packed_elems = info.packed_elems
        self.smem_bytes = info.smem_bytes
        self.ring_stages = info.ring_stages
        self.ctas_per_sm = info.ctas_per_sm
        self.n_units = info.n_units            # sweep units (LD blocks, or 1024-row tiles of blocks > 4096 rows)
        self.n_phases = info.n_phases          # sweep launches per E-step
        self.ext_elems = info.ext_elems
        self.elem_size = {0: 1, 1: 2, 2: 4, 3: 8}[info.ld_dtype]
        self.device = torch.device("cuda", torch.cuda.current_device())

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("DeviceLD has been destroyed")
        return self._h

    def block_rows(self):
        out = np.empty(self.n_blocks + 1, dtype=np.int32)
        _lib.check(_lib.lib().viprs_b200_ld_block_rows(self.handle, out.ctypes.data), "viprs_b200_ld_block_rows")
        return out

    def backward_dot(self, x, q, dq_scale=1.0):
        """q[j] += dq_scale * sum_{k>j} R_jk x[k]  (update_q_factor, e_step.hpp:307-338)."""
        L = _lib.lib()
        assert x.is_cuda and q.is_cuda and x.dtype == q.dtype and x.is_contiguous() and q.is_contiguous()
        fn = L.viprs_b200_backward_dot_f32 if x.dtype == torch.float32 else L.viprs_b200_backward_dot_f64
        _lib.check(fn(self.handle, x.data_ptr(), q.data_ptr(), float(dq_scale), _stream_ptr()), "backward_dot")

    def destroy(self):
        if getattr(self, "_h", None):
            _lib.lib().viprs_b200_ld_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
