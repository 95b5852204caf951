"""
ctypes binding of the C ABI in include/viprs_b200.h (libviprs_b200.so, built in-tree by
``__graft_entry__.build()`` / ``viprs_b200.build``).  There is no CPU fallback: if the shared
library is missing, or no CUDA device is visible, every compute entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VIPRS_B200_LIB: load an alternative build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("VIPRS_B200_LIB") or os.path.join(_HERE, "_C", "libviprs_b200.so")

I8, I16, F32, F64 = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
NSUMS = 16


class ViprsB200Error(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        super().__init__(f"viprs_b200: {what} failed with code {code}: {strerror(code)}")


class LdInfo(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int32), ("ld_dtype", ctypes.c_int32), ("n_blocks", ctypes.c_int32),
                ("max_block", ctypes.c_int32), ("n_panels", ctypes.c_int32), ("stage_bytes", ctypes.c_int32),
                ("nnz", ctypes.c_int64), ("packed_elems", ctypes.c_int64), ("smem_bytes", ctypes.c_int64),
                ("ring_stages", ctypes.c_int32), ("ctas_per_sm", ctypes.c_int32),
                ("n_units", ctypes.c_int32), ("n_phases", ctypes.c_int32), ("ext_elems", ctypes.c_int64)]


_lib = None

_vp, _i32, _i64, _f32, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> (restype, argtypes); every symbol declared in include/viprs_b200.h
SIGNATURES = {
    "viprs_b200_version": (ctypes.c_char_p, []),
    "viprs_b200_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "viprs_b200_device_count": (ctypes.c_int, []),
    "viprs_b200_ld_create": (ctypes.c_int, [ctypes.POINTER(_vp), _i32, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp]),
    "viprs_b200_ld_info": (ctypes.c_int, [_vp, ctypes.POINTER(LdInfo)]),
    "viprs_b200_ld_block_rows": (ctypes.c_int, [_vp, _vp]),
    "viprs_b200_ld_destroy": (ctypes.c_int, [_vp]),
    "viprs_b200_e_step_f32": (ctypes.c_int, [_vp] * 10 + [_f32, _i32, _vp, _vp]),
    "viprs_b200_e_step_f64": (ctypes.c_int, [_vp] * 10 + [_f64, _i32, _vp, _vp]),
    "viprs_b200_q_offset_f32": (ctypes.c_int, [_vp, _vp, _vp, _f32, _vp, _vp]),
    "viprs_b200_q_offset_f64": (ctypes.c_int, [_vp, _vp, _vp, _f64, _vp, _vp]),
    "viprs_b200_e_step_mixture_f32": (ctypes.c_int, [_vp, _i32] + [_vp] * 10 + [_f32, _i32, _vp, _vp]),
    "viprs_b200_e_step_mixture_f64": (ctypes.c_int, [_vp, _i32] + [_vp] * 10 + [_f64, _i32, _vp, _vp]),
    "viprs_b200_e_step_incremental_f32": (ctypes.c_int, [_vp] * 10 + [_f32, _vp]),
    "viprs_b200_e_step_mixture_incremental_f32": (ctypes.c_int, [_vp, _i32] + [_vp] * 10 + [_f32, _vp]),
    "viprs_b200_e_step_grid_f32": (ctypes.c_int, [_vp, _i32, _i32] + [_vp] * 10 + [_f32, _vp]),
    "viprs_b200_e_step_grid_f64": (ctypes.c_int, [_vp, _i32, _i32] + [_vp] * 10 + [_f64, _vp]),
    "viprs_b200_backward_dot_f32": (ctypes.c_int, [_vp, _vp, _vp, _f32, _vp]),
    "viprs_b200_backward_dot_f64": (ctypes.c_int, [_vp, _vp, _vp, _f64, _vp]),
    "viprs_b200_cpp_e_step": (ctypes.c_int, [_i32, _vp, _vp, _i32, _vp, _i32, _i32] + [_vp] * 9 + [_f64, _i32, _i32]),
    "viprs_b200_cpp_e_step_mixture": (ctypes.c_int, [_i32, _i32, _vp, _vp, _i32, _vp, _i32, _i32] + [_vp] * 10 + [_f64, _i32, _i32]),
    "viprs_b200_cpp_e_step_resident": (ctypes.c_int, [_vp, _i32] + [_vp] * 9 + [_f64, _i32, _vp]),
    "viprs_b200_cpp_e_step_mixture_resident": (ctypes.c_int, [_vp, _i32, _i32] + [_vp] * 10 + [_f64, _i32, _vp]),
    "viprs_b200_prepare_f32": (ctypes.c_int, [_i32, _i32, _i32, _i32] + [_vp] * 7),
    "viprs_b200_prepare_f64": (ctypes.c_int, [_i32, _i32, _i32, _i32] + [_vp] * 7),
    "viprs_b200_sums_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "viprs_b200_sums_f32": (ctypes.c_int, [_i32, _i32, _i32, _i32] + [_vp] * 10 + [_f64, _vp, _i64, _vp, _vp]),
    "viprs_b200_sums_f64": (ctypes.c_int, [_i32, _i32, _i32, _i32] + [_vp] * 10 + [_f64, _vp, _i64, _vp, _vp]),
    "viprs_b200_e_step_fused_f32": (ctypes.c_int, [_vp] * 10 + [_f32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "viprs_b200_em_update": (ctypes.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _f64, _f64, _f64, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "viprs_b200_cpp_e_step_grid": (ctypes.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _i32, _i32] + [_vp] * 9 + [_f64, _i32, _i32]),
}


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  viprs_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def strerror(code):
    try:
        return lib().viprs_b200_strerror(int(code)).decode()
    except Exception:  # pragma: no cover
        return "?"


def check(code, what):
    if code != 0:
        raise ViprsB200Error(code, what)
