"""Stage-level operators on CUDA tensors: thin, typed wrappers over the C ABI.

Every function enqueues on torch's current stream and returns torch tensors on the same
device.  Reference call sites are cited in include/b200vqa.h.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

PATCH, TARGET, TOP_N = 16, 224, 196
BILINEAR, LANCZOS = 0, 1


class Context:
    """Owns a b200vqa_t handle (weights, TMA descriptors, workspaces) on one device."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise _lib.B200VQAError("no CUDA device: relax_vqa_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        check(self.lib.b200vqa_create(device, C.byref(h)), "b200vqa_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200vqa_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.b200vqa_launch_count(self.h))

    def set_profiling(self, on):
        check(self.lib.b200vqa_set_profiling(self.h, int(on)), "set_profiling")

    def profile_read(self):
        """-> (ms spent in tcgen05 GEMM/conv launches, launches, algorithmic FLOPs) since the last read."""
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        check(self.lib.b200vqa_profile_read(self.h, C.byref(ms), C.byref(n), C.byref(fl)), "profile_read")
        return ms.value, n.value, fl.value

    def profile_read_flow(self):
        """-> (ms spent in k4_flow_iter launches, launches, algorithmic bytes) since the last read."""
        ms, n, by = C.c_double(), C.c_int64(), C.c_double()
        check(self.lib.b200vqa_profile_read_flow(self.h, C.byref(ms), C.byref(n), C.byref(by)), "profile_read_flow")
        return ms.value, n.value, by.value

    def set_flow_impl(self, impl):
        check(self.lib.b200vqa_set_flow_impl(self.h, int(impl)), "set_flow_impl")

    def set_attn_impl(self, impl):
        check(self.lib.b200vqa_set_attn_impl(self.h, int(impl)), "set_attn_impl")

    def set_gemm_sms(self, sms):
        """Persistent tcgen05 grids use at most `sms` SMs (0 = all): room for concurrent bandwidth kernels."""
        check(self.lib.b200vqa_set_gemm_sms(self.h, int(sms)), "set_gemm_sms")

    def set_gemm_impl(self, impl):
        check(self.lib.b200vqa_set_gemm_impl(self.h, int(impl)), "set_gemm_impl")


def _u8(t):
    assert t.dtype == torch.uint8 and t.is_cuda and t.is_contiguous()
    return t


def yuv420p_to_bgr(yuv, H, W):
    """yuv [B, H*W*3/2] u8 (planar yuv420p frames) -> [B,H,W,3] u8 BGR: swscale's unscaled BT.601 conversion, bit-exact."""
    lib = _lib.load()
    B = _u8(yuv).shape[0]
    assert yuv.shape[1] == H * W * 3 // 2
    bgr = torch.empty((B, H, W, 3), dtype=torch.uint8, device=yuv.device)
    check(lib.b200vqa_yuv420p_to_bgr(ptr(yuv), B, H, W, ptr(bgr), stream_ptr(yuv.device)), "yuv420p_to_bgr")
    return bgr


def absdiff_patchsum(frame, nxt, want_residual=False, want_gray=True):
    """frame/nxt [B,H,W,3] u8 BGR -> dict(sums [B,gh,gw] int32-valued uint32 bits, residual, gray0, gray1)."""
    lib = _lib.load()
    B, H, W, _ = _u8(frame).shape
    assert _u8(nxt).shape == frame.shape
    dev = frame.device
    sums = torch.empty((B, H // PATCH, W // PATCH), dtype=torch.int32, device=dev)
    residual = torch.empty_like(frame) if want_residual else None
    g0 = torch.empty((B, H, W), dtype=torch.uint8, device=dev) if want_gray else None
    g1 = torch.empty((B, H, W), dtype=torch.uint8, device=dev) if want_gray else None
    check(lib.b200vqa_absdiff_patchsum_u8(ptr(frame), ptr(nxt), B, H, W, ptr(residual), ptr(sums), ptr(g0), ptr(g1),
                                          stream_ptr(dev)), "absdiff_patchsum")
    return dict(sums=sums, residual=residual, gray0=g0, gray1=g1)


def patchsum(img):
    lib = _lib.load()
    B, H, W, _ = _u8(img).shape
    sums = torch.empty((B, H // PATCH, W // PATCH), dtype=torch.int32, device=img.device)
    check(lib.b200vqa_patchsum_u8(ptr(img), B, H, W, ptr(sums), stream_ptr(img.device)), "patchsum")
    return sums


def topk_patches(sums, top_n=TOP_N):
    """sums [B,gh,gw] -> (pos [B,top_n,2] int32 (y,x) raster order, -1 padded; count [B] int32)."""
    lib = _lib.load()
    B, gh, gw = sums.shape
    pos = torch.empty((B, top_n, 2), dtype=torch.int32, device=sums.device)
    count = torch.empty((B,), dtype=torch.int32, device=sums.device)
    check(lib.b200vqa_topk_patches(ptr(sums), B, gh, gw, top_n, ptr(pos), ptr(count), stream_ptr(sums.device)), "topk")
    return pos, count


def gather_fragments(frame, nxt, pos, count, want_ori=True, want_diff=True):
    lib = _lib.load()
    B, H, W, _ = _u8(frame).shape
    dev = frame.device
    ori = torch.empty((B, TARGET, TARGET, 3), dtype=torch.uint8, device=dev) if want_ori else None
    diff = torch.empty((B, TARGET, TARGET, 3), dtype=torch.uint8, device=dev) if want_diff else None
    check(lib.b200vqa_gather_fragments(ptr(frame), ptr(nxt) if nxt is not None else None, B, H, W, ptr(pos), ptr(count),
                                       pos.shape[1], ptr(ori), ptr(diff), stream_ptr(dev)), "gather_fragments")
    return ori, diff


def merge_fragments(a, b):
    lib = _lib.load()
    assert a.shape == b.shape
    out = torch.empty_like(_u8(a))
    check(lib.b200vqa_merge_fragments(ptr(a), ptr(_u8(b)), a.numel(), ptr(out), stream_ptr(a.device)), "merge")
    return out


def resize_pil(ctx, src, filt, swap_rb=False):
    B, H, W, _ = _u8(src).shape
    dst = torch.empty((B, TARGET, TARGET, 3), dtype=torch.uint8, device=src.device)
    check(ctx.lib.b200vqa_resize_pil(ctx.h, ptr(src), B, H, W, int(filt), int(swap_rb), ptr(dst), stream_ptr(src.device)), "resize_pil")
    return dst


def resize_pil_pair(ctx, src, swap_rb=False):
    """(BILINEAR, LANCZOS) resizes of the same frames in one call: the horizontal passes share one read of the source.
    Bit-identical to two resize_pil calls."""
    B, H, W, _ = _u8(src).shape
    bil = torch.empty((B, TARGET, TARGET, 3), dtype=torch.uint8, device=src.device)
    lan = torch.empty_like(bil)
    check(ctx.lib.b200vqa_resize_pil_pair(ctx.h, ptr(src), B, H, W, int(swap_rb), ptr(bil), ptr(lan), stream_ptr(src.device)), "resize_pil_pair")
    return bil, lan


def gemm_f16(ctx, A, Bm, bias=None, impl=0):
    """D = A @ Bm.T + bias; A [M,K], Bm [N,K] fp16 -> fp32 [M,N] (test / profiling entry)."""
    M, K = A.shape
    N = Bm.shape[0]
    D = torch.empty((M, N), dtype=torch.float32, device=A.device)
    check(ctx.lib.b200vqa_gemm_f16(ctx.h, ptr(A), ptr(Bm), ptr(bias), ptr(D), M, N, K, impl, stream_ptr(A.device)), "gemm_f16")
    return D


def farneback(ctx, gray0, gray1):
    """gray0/gray1 [B,H,W] u8 -> flow [B,H,W,2] f32 (cv2.calcOpticalFlowFarneback(...,0.5,3,15,3,5,1.2,0))."""
    B, H, W = _u8(gray0).shape
    flow = torch.empty((B, H, W, 2), dtype=torch.float32, device=gray0.device)
    check(ctx.lib.b200vqa_farneback(ctx.h, ptr(gray0), ptr(_u8(gray1)), B, H, W, ptr(flow), stream_ptr(gray0.device)), "farneback")
    return flow


def farneback_flow_sums(ctx, gray0, gray1):
    """farneback + the colour statistics of flow_to_rgb in one call -> (flow [B,H,W,2], sums [B,gh,gw], minmax [B,2]);
    the magnitude extrema come from the launch that writes the flow.  Bit-identical to farneback(); flow_to_rgb(want_rgb=False)."""
    B, H, W = _u8(gray0).shape
    dev = gray0.device
    flow = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev)
    sums = torch.empty((B, H // PATCH, W // PATCH), dtype=torch.int32, device=dev)
    minmax = torch.empty((B, 2), dtype=torch.float32, device=dev)
    check(ctx.lib.b200vqa_farneback_flow_sums(ctx.h, ptr(gray0), ptr(_u8(gray1)), B, H, W, ptr(flow), ptr(sums), ptr(minmax),
                                              stream_ptr(dev)), "farneback_flow_sums")
    return flow, sums, minmax


def flow_to_rgb(flow, want_rgb=True, want_sums=True):
    """flow [B,H,W,2] -> (rgb [B,H,W,3] u8 BGR or None, sums [B,gh,gw] or None, minmax [B,2])."""
    lib = _lib.load()
    B, H, W, _ = flow.shape
    dev = flow.device
    rgb = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) if want_rgb else None
    sums = torch.empty((B, H // PATCH, W // PATCH), dtype=torch.int32, device=dev) if want_sums else None
    minmax = torch.empty((B, 2), dtype=torch.float32, device=dev)
    check(lib.b200vqa_flow_to_rgb(ptr(flow), B, H, W, ptr(rgb), ptr(sums), ptr(minmax), stream_ptr(dev)), "flow_to_rgb")
    return rgb, sums, minmax


def flow_fragment_merge(flow, minmax, pos, count, diff_frag, want_flow_frag=False):
    lib = _lib.load()
    B, H, W, _ = flow.shape
    dev = flow.device
    flow_frag = torch.empty((B, TARGET, TARGET, 3), dtype=torch.uint8, device=dev) if want_flow_frag else None
    merged = torch.empty((B, TARGET, TARGET, 3), dtype=torch.uint8, device=dev)
    check(lib.b200vqa_flow_fragment_merge(ptr(flow), ptr(minmax), B, H, W, ptr(pos), ptr(count), pos.shape[1],
                                          ptr(diff_frag), ptr(flow_frag), ptr(merged), stream_ptr(dev)), "flow_fragment_merge")
    return flow_frag, merged


# ------------------------------------------------------------------ networks and head
def _load_state_dict(ctx, fn, sd, what):
    items = [(k, v.detach().to(torch.float32).contiguous().cpu()) for k, v in sd.items()
             if torch.is_tensor(v) and v.dtype.is_floating_point]
    n = len(items)
    names = (C.c_char_p * n)(*[k.encode() for k, _ in items])
    ptrs = (C.c_void_p * n)(*[v.data_ptr() for _, v in items])
    numels = (C.c_int64 * n)(*[v.numel() for _, v in items])
    check(fn(ctx.h, n, names, ptrs, numels), what)


def load_resnet50(ctx, state_dict):
    """state_dict: torchvision resnet50 keys (pretrained=True in the reference, visualise_resnet.py:21)."""
    _load_state_dict(ctx, ctx.lib.b200vqa_load_resnet50, state_dict, "load_resnet50")


def load_vitb16(ctx, state_dict):
    """state_dict: DINO ViT-B/16 keys (visualise_vit_layer.py:304-330)."""
    _load_state_dict(ctx, ctx.lib.b200vqa_load_vitb16, state_dict, "load_vitb16")


def load_head(ctx, state_dict, imputer_mean, scaler_scale, scaler_min):
    """Mlp state dict (possibly SWA-wrapped) + fitted imputer / scaler attributes (float64)."""
    from .weights import fix_state_dict
    import numpy as np
    sd = {k: v.detach().to(torch.float32).contiguous().cpu() for k, v in fix_state_dict(state_dict).items()
          if torch.is_tensor(v) and v.dtype.is_floating_point}
    keep = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in (imputer_mean, scaler_scale, scaler_min)]
    p = lambda t: C.c_void_p(t.data_ptr())
    check(ctx.lib.b200vqa_load_head(ctx.h, sd["fc1.weight"].shape[1], p(sd["fc1.weight"]), p(sd["fc1.bias"]), p(sd["bn1.weight"]),
                                    p(sd["bn1.bias"]), p(sd["bn1.running_mean"]), p(sd["bn1.running_var"]), p(sd["fc2.weight"]),
                                    p(sd["fc2.bias"]), p(sd["fc3.weight"]), p(sd["fc3.bias"]),
                                    C.c_void_p(keep[0].ctypes.data), C.c_void_p(keep[1].ctypes.data), C.c_void_p(keep[2].ctypes.data)),
          "load_head")


def resnet50_features(ctx, img, is_bgr=True, want_stack=True, want_pool=False):
    """img [B,224,224,3] u8 -> (stack [B,13120] | None, pool [B,2051] | None)."""
    B = _u8(img).shape[0]
    stack = torch.empty((B, 13120), dtype=torch.float32, device=img.device) if want_stack else None
    pool = torch.empty((B, 2051), dtype=torch.float32, device=img.device) if want_pool else None
    check(ctx.lib.b200vqa_resnet50_features(ctx.h, ptr(img), B, int(is_bgr), ptr(stack), ptr(pool), stream_ptr(img.device)),
          "resnet50_features")
    return stack, pool


def vitb16_features(ctx, img, is_bgr=True):
    B = _u8(img).shape[0]
    out = torch.empty((B, 2304), dtype=torch.float32, device=img.device)
    check(ctx.lib.b200vqa_vitb16_features(ctx.h, ptr(img), B, int(is_bgr), ptr(out), stream_ptr(img.device)), "vitb16_features")
    return out


RESNET_MAP_SHAPES = [(64, 112, 112)] + [(256, 56, 56)] * 3 + [(512, 28, 28)] * 4 + [(1024, 14, 14)] * 4 + [(2048, 7, 7)] * 3
RESNET_MAP_FLOATS = sum(c * h * w for c, h, w in RESNET_MAP_SHAPES)


def resnet50_maps(ctx, img, is_bgr=True):
    """img [B,224,224,3] u8 -> list of 15 fp32 tensors [B,C,H,W]: the raw hooked activations the reference's
    visualise_resnet.process_video_frame returns (boundary-fidelity path; the fast path pools in the conv epilogues)."""
    B = _u8(img).shape[0]
    flat = torch.empty((B, RESNET_MAP_FLOATS), dtype=torch.float32, device=img.device)
    check(ctx.lib.b200vqa_resnet50_maps(ctx.h, ptr(img), B, int(is_bgr), ptr(flat), stream_ptr(img.device)), "resnet50_maps")
    out, o = [], 0
    for c, h, w in RESNET_MAP_SHAPES:
        out.append(flat[:, o:o + c * h * w].reshape(B, c, h, w))
        o += c * h * w
    return out


def vitb16_tokens(ctx, img, is_bgr=True):
    """img [B,224,224,3] u8 -> final-LayerNorm patch tokens [B,196,768] fp32 (norm(x)[:, 1:], visualise_vit_layer.py:234-239)."""
    B = _u8(img).shape[0]
    out = torch.empty((B, 196, 768), dtype=torch.float32, device=img.device)
    check(ctx.lib.b200vqa_vitb16_tokens(ctx.h, ptr(img), B, int(is_bgr), ptr(out), stream_ptr(img.device)), "vitb16_tokens")
    return out


def temporal_mean_concat(full_stack, full_vit, frag_stack, frag_pool, frag_vit_ori, frag_vit_mer, full_off, pair_off):
    lib = _lib.load()
    V = full_off.numel() - 1
    feats = torch.empty((V, 35203), dtype=torch.float32, device=full_stack.device)
    check(lib.b200vqa_temporal_mean_concat(ptr(full_stack), ptr(full_vit), ptr(frag_stack), ptr(frag_pool), ptr(frag_vit_ori),
                                           ptr(frag_vit_mer), ptr(full_off), ptr(pair_off), V, ptr(feats),
                                           stream_ptr(full_stack.device)), "temporal_mean_concat")
    return feats


def head_forward(ctx, features):
    V = features.shape[0]
    score = torch.empty((V,), dtype=torch.float32, device=features.device)
    check(ctx.lib.b200vqa_head_forward(ctx.h, ptr(features), V, ptr(score), stream_ptr(features.device)), "head_forward")
    return score
