"""
Drop-in replacements for the reference's Cython entry points
(/root/reference/viprs/model/vi/e_step_cpp.pyx): same names, same positional arguments, same
in-place outputs.  numpy arguments take the one-shot host path of the C ABI
(upload + sweep + download inside the call); torch CUDA tensors take the device path and re-use a
cached device-resident LD matrix.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib
from .ld import DeviceLD, _NP_DT, _stream_ptr

_FLOAT_DT = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64}
_ld_cache = {}            # insertion-ordered: (DeviceLD, tensor versions, weakrefs)
_LD_CACHE_MAX = 4


def check_omp_support():
    """e_step_cpp.pyx:75-76 -- there is no OpenMP path here; kept for API compatibility."""
    return False


def check_blas_support():
    """e_step_cpp.pyx:71-72 -- there is no BLAS path here; kept for API compatibility."""
    return False


def device_ld_for(ld_left_bound, ld_indptr, ld_data):
    """
    Device LD cache for callers that pass the same CUDA tensors every iteration (the cpp_e_step* drop-ins with
    torch arguments).  An entry is keyed on the identity of the three tensor OBJECTS, stays valid only while their
    in-place version counters are unchanged, and is evicted when any of them is garbage-collected -- a freed and
    re-allocated buffer at the same address can therefore never pick up a stale LD matrix.  At most
    ``_LD_CACHE_MAX`` repacked copies are kept alive (least recently used goes first).
    """
    tensors = (ld_data, ld_indptr, ld_left_bound)
    key = tuple(id(t) for t in tensors)
    versions = tuple(t._version for t in tensors)
    ent = _ld_cache.get(key)
    if ent is not None and ent[1] == versions and all(r() is t for r, t in zip(ent[2], tensors)):
        _ld_cache[key] = _ld_cache.pop(key)                  # most recently used last
        return ent[0]
    ld = DeviceLD(ld_data, ld_indptr, ld_left_bound)
    refs = tuple(weakref.ref(t, lambda _r, k=key: _ld_cache.pop(k, None)) for t in tensors)
    _ld_cache.pop(key, None)
    _ld_cache[key] = (ld, versions, refs)
    while len(_ld_cache) > _LD_CACHE_MAX:
        _ld_cache.pop(next(iter(_ld_cache)))
    return ld


def _opt_ptr(t, ld, dt, what):
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous() and t.dtype == dt and t.numel() == ld.M):
        raise ValueError(f"{what}: q_offset must be a contiguous CUDA tensor of length M and the state dtype")
    return t.data_ptr()


def q_offset_device(ld, eta, q, dq_scale, out=None):
    """
    ``q - dq_scale * (R - I) eta``: the part of a caller's q that eta does not explain.  The reference keeps q
    incrementally (e_step.hpp:421,439), so this part survives every sweep; pass the result as ``q_offset`` to
    ``e_step_device`` / ``e_step_mixture_device`` to reproduce that (e.g. after a ``param_0`` warm start, where
    eta != 0 next to q = 0, VIPRS.py:339-357).
    """
    L = _lib.lib()
    out = torch.empty_like(q) if out is None else out
    for t in (eta, q, out):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == q.dtype and t.numel() == ld.M):
            raise ValueError("q_offset_device: arrays must be contiguous CUDA tensors of length M and one dtype")
    fn = L.viprs_b200_q_offset_f32 if q.dtype == torch.float32 else L.viprs_b200_q_offset_f64
    _lib.check(fn(ld.handle, eta.data_ptr(), q.data_ptr(), float(dq_scale), out.data_ptr(), _stream_ptr()), "viprs_b200_q_offset")
    return out


def e_step_device(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult,
                  dq_scale, materialize_q=True, q_offset=None):
    """One sweep on a DeviceLD; all arrays are contiguous torch CUDA tensors of one float dtype."""
    L = _lib.lib()
    ts = (std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult)
    dt = var_mu.dtype
    for t in ts:
        if not (t.is_cuda and t.is_contiguous() and t.dtype == dt and t.numel() == ld.M):
            raise ValueError("e_step_device: arrays must be contiguous CUDA tensors of length M and one dtype")
    fn = L.viprs_b200_e_step_f32 if dt == torch.float32 else L.viprs_b200_e_step_f64
    rc = fn(ld.handle, *[t.data_ptr() for t in ts], float(dq_scale), int(bool(materialize_q)),
            _opt_ptr(q_offset, ld, dt, "e_step_device"), _stream_ptr())
    _lib.check(rc, "viprs_b200_e_step")


def e_step_incremental_device(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult,
                              dq_scale):
    """
    cpp_e_step (e_step_cpp.pyx:91-122) on float32 CUDA tensors with the reference's own incremental bookkeeping of q
    (e_step.hpp:421, 435-440): q is in/out, whatever it holds on entry is honoured.  Raises ViprsB200Error
    (VIPRS_B200_EUNSUPPORTED) where the register-resident sweep does not apply.
    """
    L = _lib.lib()
    ts = (std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult)
    for t in ts:
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == ld.M):
            raise ValueError("e_step_incremental_device: arrays must be contiguous float32 CUDA tensors of length M")
    rc = L.viprs_b200_e_step_incremental_f32(ld.handle, *[t.data_ptr() for t in ts], float(dq_scale), _stream_ptr())
    _lib.check(rc, "viprs_b200_e_step_incremental_f32")


def e_step_mixture_incremental_device(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs,
                                      sqrt_half_var_tau, mu_mult, dq_scale):
    """cpp_e_step_mixture (e_step_cpp.pyx:125-159) on float32 CUDA tensors, q in/out and incremental (K <= 4)."""
    L = _lib.lib()
    if var_mu.dim() != 2:
        raise ValueError("e_step_mixture_incremental_device: var_mu must be (M, K)")
    K = var_mu.shape[1]
    for t in (var_gamma, var_mu, u_logs, sqrt_half_var_tau, mu_mult):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and tuple(t.shape) == (ld.M, K)):
            raise ValueError("e_step_mixture_incremental_device: (M,K) arrays must be C-contiguous float32 CUDA tensors")
    for t in (std_beta, eta, q, eta_diff, log_null_pi):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == ld.M):
            raise ValueError("e_step_mixture_incremental_device: (M,) arrays must be contiguous float32 CUDA tensors")
    rc = L.viprs_b200_e_step_mixture_incremental_f32(
        ld.handle, K, std_beta.data_ptr(), var_gamma.data_ptr(), var_mu.data_ptr(), eta.data_ptr(), q.data_ptr(),
        eta_diff.data_ptr(), log_null_pi.data_ptr(), u_logs.data_ptr(), sqrt_half_var_tau.data_ptr(), mu_mult.data_ptr(),
        float(dq_scale), _stream_ptr())
    _lib.check(rc, "viprs_b200_e_step_mixture_incremental_f32")


def e_step_mixture_device(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs,
                          sqrt_half_var_tau, mu_mult, dq_scale, materialize_q=True, q_offset=None):
    """One mixture sweep on a DeviceLD; (M,K) arrays are C-order CUDA tensors, the rest have M entries."""
    L = _lib.lib()
    dt = var_mu.dtype
    if var_mu.dim() != 2:
        raise ValueError("e_step_mixture_device: var_mu must be (M, K)")
    K = var_mu.shape[1]
    for t in (var_gamma, var_mu, u_logs, sqrt_half_var_tau, mu_mult):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == dt and tuple(t.shape) == (ld.M, K)):
            raise ValueError("e_step_mixture_device: (M,K) arrays must be C-contiguous CUDA tensors of one dtype")
    for t in (std_beta, eta, q, eta_diff, log_null_pi):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == dt and t.numel() == ld.M):
            raise ValueError("e_step_mixture_device: (M,) arrays must be contiguous CUDA tensors of one dtype")
    fn = L.viprs_b200_e_step_mixture_f32 if dt == torch.float32 else L.viprs_b200_e_step_mixture_f64
    rc = fn(ld.handle, K, std_beta.data_ptr(), var_gamma.data_ptr(), var_mu.data_ptr(), eta.data_ptr(), q.data_ptr(),
            eta_diff.data_ptr(), log_null_pi.data_ptr(), u_logs.data_ptr(), sqrt_half_var_tau.data_ptr(),
            mu_mult.data_ptr(), float(dq_scale), int(bool(materialize_q)),
            _opt_ptr(q_offset, ld, dt, "e_step_mixture_device"), _stream_ptr())
    _lib.check(rc, "viprs_b200_e_step_mixture")


def e_step_grid_device(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, half_var_tau, mu_mult, dq_scale,
                       active_model_idx):
    """
    One grid sweep on a DeviceLD.  (M,G) arrays are COLUMN-MAJOR CUDA tensors -- i.e. torch tensors of shape
    (G, M) that are C-contiguous, or any (M, G) tensor ``t`` with ``t.t().is_contiguous()``.  q is in/out.
    ``active_model_idx``: int32 CUDA tensor of the columns to update.
    """
    L = _lib.lib()
    dt = var_mu.dtype
    mats = (var_gamma, var_mu, eta, q, eta_diff, u_logs, half_var_tau, mu_mult)

    def colmajor(t):
        if t.dim() != 2:
            raise ValueError("e_step_grid_device: (M,G) arrays must be 2-D")
        if t.shape[0] == ld.M and t.t().is_contiguous():
            return t.shape[1]
        raise ValueError("e_step_grid_device: (M,G) arrays must be column-major (Fortran order) with M rows")

    G = colmajor(var_mu)
    for t in mats:
        if not (t.is_cuda and t.dtype == dt and colmajor(t) == G):
            raise ValueError("e_step_grid_device: arrays must be CUDA tensors of one float dtype and one shape")
    if not (std_beta.is_cuda and std_beta.dtype == dt and std_beta.is_contiguous() and std_beta.numel() == ld.M):
        raise ValueError("e_step_grid_device: std_beta must be a contiguous CUDA tensor of length M")
    act = active_model_idx
    if not (act.is_cuda and act.dtype == torch.int32 and act.is_contiguous()):
        raise ValueError("e_step_grid_device: active_model_idx must be a contiguous int32 CUDA tensor")
    fn = L.viprs_b200_e_step_grid_f32 if dt == torch.float32 else L.viprs_b200_e_step_grid_f64
    rc = fn(ld.handle, G, act.numel(), act.data_ptr(), std_beta.data_ptr(), *[t.data_ptr() for t in mats],
            float(dq_scale), _stream_ptr())
    _lib.check(rc, "viprs_b200_e_step_grid")


def cpp_e_step_grid(ld_left_bound, ld_indptr, ld_data, std_beta, var_gamma, var_mu, eta, q, eta_diff,
                    u_logs, half_var_tau, mu_mult, dq_scale, active_model_idx, threads=1, low_memory=True):
    """
    cpp_e_step_grid (e_step_cpp.pyx:161-195): (M,G) arrays Fortran-order, in-place outputs, only the columns in
    ``active_model_idx`` are touched.  ``threads`` is accepted and ignored (always the sequential sweep).
    """
    if isinstance(ld_data, torch.Tensor):
        ld = device_ld_for(ld_left_bound, ld_indptr, ld_data)
        act = torch.as_tensor(active_model_idx, dtype=torch.int32, device=ld_data.device).contiguous()
        return e_step_grid_device(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, half_var_tau, mu_mult,
                                  dq_scale, act)
    L = _lib.lib()
    M, G = var_mu.shape
    dt = var_mu.dtype
    fdt = _FLOAT_DT[dt]
    lb, ip, ld_data = _host_index_arrays(ld_left_bound, ld_indptr, ld_data)
    act = np.ascontiguousarray(active_model_idx, dtype=np.int32)
    if act.size and (act.min() < 0 or act.max() >= G):
        raise ValueError("cpp_e_step_grid: active_model_idx out of range")
    beta = np.ascontiguousarray(std_beta, dtype=dt)
    ins = []
    for a in (u_logs, half_var_tau, mu_mult):
        a = np.asfortranarray(a, dtype=dt)
        if a.shape != (M, G):
            raise ValueError("cpp_e_step_grid: u_logs / half_var_tau / mu_mult must be (M, G)")
        ins.append(a)
    for a in (var_gamma, var_mu, eta, q, eta_diff):
        if not (a.flags["F_CONTIGUOUS"] and a.dtype == dt and a.shape == (M, G)):
            raise ValueError("cpp_e_step_grid: in/out arrays must be Fortran-contiguous (M, G) of one float dtype")
    rc = L.viprs_b200_cpp_e_step_grid(M, G, act.shape[0], act.ctypes.data, lb.ctypes.data, ip.ctypes.data,
                                      int(ip.dtype == np.int64), ld_data.ctypes.data, _NP_DT[ld_data.dtype], fdt,
                                      beta.ctypes.data, var_gamma.ctypes.data, var_mu.ctypes.data, eta.ctypes.data,
                                      q.ctypes.data, eta_diff.ctypes.data, ins[0].ctypes.data, ins[1].ctypes.data,
                                      ins[2].ctypes.data, float(dq_scale), int(threads), int(bool(low_memory)))
    _lib.check(rc, "viprs_b200_cpp_e_step_grid")


def _host_index_arrays(ld_left_bound, ld_indptr, ld_data):
    lb = np.ascontiguousarray(ld_left_bound, dtype=np.int32)
    ip = np.ascontiguousarray(ld_indptr)
    if ip.dtype not in (np.int32, np.int64):
        ip = ip.astype(np.int64)
    return lb, ip, np.ascontiguousarray(ld_data)


def cpp_e_step_mixture(ld_left_bound, ld_indptr, ld_data, std_beta, var_gamma, var_mu, eta, q, eta_diff,
                       log_null_pi, u_logs, sqrt_half_var_tau, mu_mult, dq_scale, threads=1, low_memory=True):
    """
    cpp_e_step_mixture (e_step_cpp.pyx:125-159): (M,K) arrays C-order, K = var_mu.shape[1]; in-place outputs.
    ``threads`` is accepted and ignored (always the sequential sweep).
    """
    if isinstance(ld_data, torch.Tensor):
        ld = device_ld_for(ld_left_bound, ld_indptr, ld_data)
        off = q_offset_device(ld, eta, q, dq_scale)           # q is in/out like the reference's
        return e_step_mixture_device(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs,
                                     sqrt_half_var_tau, mu_mult, dq_scale, True, off)
    L = _lib.lib()
    M, K = var_mu.shape
    dt = var_mu.dtype
    fdt = _FLOAT_DT[dt]
    lb, ip, ld_data = _host_index_arrays(ld_left_bound, ld_indptr, ld_data)
    ins = [np.ascontiguousarray(a, dtype=dt) for a in (std_beta, log_null_pi, u_logs, sqrt_half_var_tau, mu_mult)]
    for a in (var_gamma, var_mu):
        if not (a.flags["C_CONTIGUOUS"] and a.dtype == dt and a.shape == (M, K)):
            raise ValueError("cpp_e_step_mixture: var_gamma / var_mu must be C-contiguous (M, K)")
    for a in (eta, q, eta_diff):
        if not (a.flags["C_CONTIGUOUS"] and a.dtype == dt and a.shape == (M,)):
            raise ValueError("cpp_e_step_mixture: eta / q / eta_diff must be C-contiguous (M,)")
    rc = L.viprs_b200_cpp_e_step_mixture(M, K, lb.ctypes.data, ip.ctypes.data, int(ip.dtype == np.int64),
                                         ld_data.ctypes.data, _NP_DT[ld_data.dtype], fdt, ins[0].ctypes.data,
                                         var_gamma.ctypes.data, var_mu.ctypes.data, eta.ctypes.data, q.ctypes.data,
                                         eta_diff.ctypes.data, ins[1].ctypes.data, ins[2].ctypes.data,
                                         ins[3].ctypes.data, ins[4].ctypes.data, float(dq_scale), int(threads),
                                         int(bool(low_memory)))
    _lib.check(rc, "viprs_b200_cpp_e_step_mixture")


def cpp_e_step(ld_left_bound, ld_indptr, ld_data, std_beta, var_gamma, var_mu, eta, q, eta_diff,
               u_logs, sqrt_half_var_tau, mu_mult, dq_scale, threads=1, low_memory=True):
    """
    cpp_e_step (e_step_cpp.pyx:91-122).  ``threads`` is accepted and ignored: the result is always the
    strictly sequential (threads=1) sweep.  ``q`` is in/out like the reference's: whatever it holds beyond
    dq_scale (R - I) eta on entry is carried through; on exit it equals the reference's q after update_q_factor.
    """
    if isinstance(ld_data, torch.Tensor):
        ld = device_ld_for(ld_left_bound, ld_indptr, ld_data)
        off = q_offset_device(ld, eta, q, dq_scale)           # q is in/out like the reference's
        return e_step_device(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau,
                             mu_mult, dq_scale, True, off)
    L = _lib.lib()
    M = var_mu.shape[0]
    fdt = _FLOAT_DT[var_mu.dtype]
    lb = np.ascontiguousarray(ld_left_bound, dtype=np.int32)
    ip = np.ascontiguousarray(ld_indptr)
    if ip.dtype not in (np.int32, np.int64):
        ip = ip.astype(np.int64)
    ins = [np.ascontiguousarray(a, dtype=var_mu.dtype) for a in (std_beta, u_logs, sqrt_half_var_tau, mu_mult)]
    for a in (var_gamma, var_mu, eta, q, eta_diff):
        if not (a.flags["C_CONTIGUOUS"] and a.dtype == var_mu.dtype and a.shape[0] == M):
            raise ValueError("cpp_e_step: in/out arrays must be C-contiguous, length M, one float dtype")
    ld_data = np.ascontiguousarray(ld_data)
    rc = L.viprs_b200_cpp_e_step(M, lb.ctypes.data, ip.ctypes.data, int(ip.dtype == np.int64),
                                 ld_data.ctypes.data, _NP_DT[ld_data.dtype], fdt, ins[0].ctypes.data,
                                 var_gamma.ctypes.data, var_mu.ctypes.data, eta.ctypes.data, q.ctypes.data,
                                 eta_diff.ctypes.data, ins[1].ctypes.data, ins[2].ctypes.data, ins[3].ctypes.data,
                                 float(dq_scale), int(threads), int(bool(low_memory)))
    _lib.check(rc, "viprs_b200_cpp_e_step")


def _host_ptr(a):
    """Host address of a numpy array or a CPU torch tensor (pinned tensors make the copies asynchronous)."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda or not a.is_contiguous():
            raise ValueError("host arrays must be contiguous CPU tensors or numpy arrays")
        return a.data_ptr()
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("host arrays must be C-contiguous")
    return a.ctypes.data


def cpp_e_step_resident(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult, dq_scale,
                        q_is_consistent=False):
    """
    cpp_e_step (e_step_cpp.pyx:91-122) for a caller that keeps the LD matrix resident on the device (``ld``: a
    ``DeviceLD``) and its state in HOST memory like the reference: one C-ABI call uploads what cpp_e_step reads, sweeps,
    materialises q and downloads what it writes.  ``q_is_consistent=True`` skips the two LD passes that carry an
    unexplained q_in (see ``q_offset_device``).
    """
    L = _lib.lib()
    dt = var_mu.dtype
    fdt = _lib.F32 if dt in (np.float32, torch.float32) else _lib.F64
    arrs = (std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult)
    for a in arrs:
        if a.dtype != dt or a.shape[0] != ld.M:
            raise ValueError("cpp_e_step_resident: arrays must have M entries and one float dtype")
    rc = L.viprs_b200_cpp_e_step_resident(ld.handle, fdt, *[_host_ptr(a) for a in arrs], float(dq_scale),
                                          int(bool(q_is_consistent)), _stream_ptr())
    _lib.check(rc, "viprs_b200_cpp_e_step_resident")


def cpp_e_step_mixture_resident(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs,
                                sqrt_half_var_tau, mu_mult, dq_scale, q_is_consistent=False):
    """cpp_e_step_mixture (e_step_cpp.pyx:125-159) with a resident LD matrix and HOST state arrays."""
    L = _lib.lib()
    dt = var_mu.dtype
    fdt = _lib.F32 if dt in (np.float32, torch.float32) else _lib.F64
    K = var_mu.shape[1]
    arrs = (std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs, sqrt_half_var_tau, mu_mult)
    rc = L.viprs_b200_cpp_e_step_mixture_resident(ld.handle, K, fdt, *[_host_ptr(a) for a in arrs], float(dq_scale),
                                                  int(bool(q_is_consistent)), _stream_ptr())
    _lib.check(rc, "viprs_b200_cpp_e_step_mixture_resident")
