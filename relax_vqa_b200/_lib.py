"""ctypes binding of libb200vqa.so (the C ABI in include/b200vqa.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that
fails, or no CUDA device is present when a context is created, an exception is raised.
PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200vqa.so")
_lib = None

c_void_p, c_int, c_size_t, c_int64 = C.c_void_p, C.c_int, C.c_size_t, C.c_int64

# name -> (restype, argtypes); mirrors include/b200vqa.h one to one
SIGNATURES = {
    "b200vqa_version": (c_int, []),
    "b200vqa_error_string": (C.c_char_p, [c_int]),
    "b200vqa_last_error": (C.c_char_p, []),
    "b200vqa_create": (c_int, [c_int, C.POINTER(c_void_p)]),
    "b200vqa_destroy": (c_int, [c_void_p]),
    "b200vqa_yuv420p_to_bgr": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200vqa_absdiff_patchsum_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200vqa_patchsum_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200vqa_topk_patches": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b200vqa_gather_fragments": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200vqa_merge_fragments": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "b200vqa_resize_pil": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200vqa_resize_pil_pair": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b200vqa_farneback": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200vqa_farneback_flow_sums": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200vqa_flow_to_rgb": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200vqa_flow_fragment_merge": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200vqa_load_resnet50": (c_int, [c_void_p, c_int, C.POINTER(C.c_char_p), C.POINTER(c_void_p), C.POINTER(c_int64)]),
    "b200vqa_load_vitb16": (c_int, [c_void_p, c_int, C.POINTER(C.c_char_p), C.POINTER(c_void_p), C.POINTER(c_int64)]),
    "b200vqa_load_head": (c_int, [c_void_p, c_int] + [c_void_p] * 13),
    "b200vqa_resnet50_features": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b200vqa_vitb16_features": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "b200vqa_resnet50_maps": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "b200vqa_vitb16_tokens": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "b200vqa_temporal_mean_concat": (c_int, [c_void_p] * 8 + [c_int, c_void_p, c_void_p]),
    "b200vqa_head_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b200vqa_trainer_create": (c_int, [c_void_p, c_int, c_int, C.POINTER(c_void_p)]),
    "b200vqa_trainer_destroy": (c_int, [c_void_p]),
    "b200vqa_trainer_set_params": (c_int, [c_void_p] + [c_void_p] * 10),
    "b200vqa_trainer_get_params": (c_int, [c_void_p, c_int] + [c_void_p] * 10),
    "b200vqa_trainer_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p] + [C.c_float] * 6 + [c_void_p, c_void_p]),
    "b200vqa_trainer_swa_update": (c_int, [c_void_p, c_void_p]),
    "b200vqa_trainer_predict": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "b200vqa_trainer_update_bn": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "b200vqa_gemm_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200vqa_launch_count": (c_int64, [c_void_p]),
    "b200vqa_set_gemm_impl": (c_int, [c_void_p, c_int]),
    "b200vqa_set_attn_impl": (c_int, [c_void_p, c_int]),
    "b200vqa_set_gemm_sms": (c_int, [c_void_p, c_int]),
    "b200vqa_set_profiling": (c_int, [c_void_p, c_int]),
    "b200vqa_set_flow_impl": (c_int, [c_void_p, c_int]),
    "b200vqa_profile_read": (c_int, [c_void_p, C.POINTER(C.c_double), C.POINTER(c_int64), C.POINTER(C.c_double)]),
    "b200vqa_profile_read_flow": (c_int, [c_void_p, C.POINTER(C.c_double), C.POINTER(c_int64), C.POINTER(C.c_double)]),
}


class B200VQAError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load (building first if needed) the shared library; raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise B200VQAError(f"{LIB_PATH} is missing and there is no CPU fallback; run __graft_entry__.build()")
        from .csrc import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header / library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        lib = load()
        msg = lib.b200vqa_error_string(rc).decode()
        detail = lib.b200vqa_last_error().decode()
        raise B200VQAError(f"{what}: {msg} ({rc}) {detail}")


def ptr(t):
    """Device pointer of a torch CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise B200VQAError("libb200vqa works on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise B200VQAError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
