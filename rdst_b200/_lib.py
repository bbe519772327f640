"""ctypes binding of librdst_b200.so (the C ABI declared in include/rdst_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a kernel call
fails, a RuntimeError is raised.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDST_B200_LIB") or os.path.join(_HERE, "lib", "librdst_b200.so")   # (override: probe builds)

F32, BF16 = 0, 1
ABI_VERSION = 1

_vp, _i64, _i, _f = C.c_void_p, C.c_int64, C.c_int, C.c_float

class PackDesc(C.Structure):
    """RdstPackDesc of include/rdst_b200.h."""
    _fields_ = [(n, C.c_void_p) for n in ("W", "b", "gamma", "beta", "Wp", "bp", "dWp", "dbp", "dW", "db", "dgamma", "dbeta")] + \
               [(n, C.c_int) for n in ("N", "K", "ldp", "scatter_rows", "scatter_cols", "q_rows")] + \
               [("q_scale", C.c_float), ("_pad", C.c_int)]


_PACK_PTR_FIELDS = ("W", "b", "gamma", "beta", "Wp", "bp", "dWp", "dbp", "dW", "db", "dgamma", "dbeta")


def pack_desc_array(descs):
    """list of dicts (tensors / None under the pointer fields of RdstPackDesc, ints and a float otherwise) -> ctypes array."""
    arr = (PackDesc * len(descs))()
    for a, d in zip(arr, descs):
        for k, v in d.items():
            setattr(a, k, (None if v is None else v.data_ptr()) if k in _PACK_PTR_FIELDS else v)
    return arr


_SIGNATURES = {
    "rdst_abi_version": (C.c_int, []),
    "rdst_last_error": (C.c_char_p, []),
    "rdst_has_tcgen05": (C.c_int, []),
    "rdst_linear_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _f, _i, _vp]),
    "rdst_window_attention_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rdst_conv3x3_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "rdst_conv3x3_act_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rdst_head_fwd": (C.c_int, [_vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp]),
    "rdst_layernorm_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _i, _f, _i, _vp]),
    "rdst_last_conv_fwd": (C.c_int, [_vp, _i64, _vp, _f, _f, _f, _vp, _i, _i, _i, _i, _i, _vp]),
    "rdst_gemm_tn_acc": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rdst_lnhat_fwd": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp]),
    "rdst_lnhat_bwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp]),
    "rdst_axpy": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i, _f, _vp]),
    "rdst_pixel_unshuffle2": (C.c_int, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp]),
    "rdst_layernorm_bwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _i, _f, _vp]),
    "rdst_gelu_fwd": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i, _vp]),
    "rdst_gelu_bwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _vp]),
    "rdst_window_attention_bwd": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "rdst_pack_linear_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "rdst_pack_linear_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "rdst_window_attention_tc_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _vp]),
    "rdst_window_attention_tc_bwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i, _i, _i, _i, _i, _vp]),
    "rdst_pack_linear_batch": (C.c_int, [_vp, _i, _i, _vp]),
    "rdst_gemm_tc": (C.c_int, [_vp, _i64, _vp, _i64, _i, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _f,
                               _i, _i, _i, _i, _i, _i, _vp]),
    "rdst_gemm_tc_lnbwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _f, _vp]),
    "rdst_gemm_tn_tc": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rdst_stl_mlp_fwd_bf16": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    "rdst_stl_attn_fwd_bf16": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rdst_conv3x3_fwd_bf16_tc": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "rdst_last_conv_fwd_bf16_tc": (C.c_int, [_vp, _i64, _vp, _f, _f, _f, _vp, _i, _i, _i, _vp]),
    "rdst_stl_mlp_tail_fwd_bf16": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _f, _i64, _i, _i, _vp]),
    "rdst_bicubic_resize_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rdst_sqdiff_sum_f64": (C.c_int, [_vp, _vp, _vp, _i, _i64, _vp]),
    "rdst_ssim_sum_f64": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "rdst_debug_attn_timing": (C.c_int, [_vp]),
    "rdst_debug_attn_variant": (C.c_int, [_i]),
    "rdst_debug_attn2_timing": (C.c_int, [_vp]),
    "rdst_debug_mlp_timing": (C.c_int, [_vp]),
    "rdst_debug_mlp_variant": (C.c_int, [_i]),
    "rdst_debug_mlp2_timing": (C.c_int, [_vp]),
    "rdst_debug_conv_timing": (C.c_int, [_vp]),
    "rdst_umma_bench": (C.c_int, [_i, _i, _i, _i, _i, _vp, _vp]),
    "rdst_tmem_bw_bench": (C.c_int, [_i, _i, _i, _vp, _vp]),
    "rdst_umma_selftest": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "rdst_tma_selftest": (C.c_int, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}

_lib = None


def exported_symbols():
    """Names every build of the library must export (checked by tests/test_abi.py against the header)."""
    return sorted(_SIGNATURES)


def load():
    """Load librdst_b200.so once; raise (never fall back) if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"rdst_b200: CUDA library not found at {LIB_PATH}. Build it with `python -m rdst_b200.build` "
            "(needs nvcc); there is no CPU or PyTorch fallback for this package.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch, which must be loud
        fn.restype, fn.argtypes = res, args
    v = lib.rdst_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"rdst_b200: ABI version mismatch: library {v}, python binding {ABI_VERSION}")
    _lib = lib
    return lib


def dtype_code(t):
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise TypeError(f"rdst_b200: unsupported activation dtype {t}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what):
    if rc != 0:
        msg = load().rdst_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"rdst_b200: {what} failed (code {rc}): {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args), name)
