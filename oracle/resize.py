"""Oracle (CPU, numpy) for A9: Pillow's 8-bit two-pass resampler, bit-exact.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Third-party dependency: pillow==10.2.0 (requirements.txt:80).  Call sites:
  ResNet path: transforms.Resize((224,224)) on a PIL image == Image.resize(BILINEAR, reducing_gap=None)
               src/extractor/visualise_resnet.py:40-47
  ViT path:    img.resize((224,224), Image.Resampling.LANCZOS)
               src/extractor/visualise_vit_layer.py:466-470
Restates Pillow's src/libImaging/Resample.c (precompute_coeffs / normalize_coeffs_8bpc /
ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc) as in SURVEY.md 8(a) A9.
"""
import math
import numpy as np

PRECISION_BITS = 32 - 8 - 2   # 22
BILINEAR, LANCZOS = 0, 1


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x *= math.pi
    return math.sin(x) / x


def _lanczos(x):
    return _sinc(x) * _sinc(x / 3.0) if -3.0 <= x < 3.0 else 0.0


_FILTERS = {BILINEAR: (_bilinear, 1.0), LANCZOS: (_lanczos, 3.0)}


def precompute_coeffs(in_size, out_size, filt):
    """-> (bounds[out,2] = (xmin, count), kk[out, ksize] int32 fixed-point weights, ksize)."""
    fn, fsup = _FILTERS[filt]
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsup * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [fn((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)            # left-to-right double accumulation, as in C
        # C accumulates ww in a loop; python's sum() does the same left fold for floats
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis1(img, out_size, filt):
    """Resample along axis 1 of (rows, in, C) uint8 -> (rows, out, C) uint8."""
    in_size = img.shape[1]
    bounds, kk, _ = precompute_coeffs(in_size, out_size, filt)
    out = np.empty((img.shape[0], out_size, img.shape[2]), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.full((img.shape[0], img.shape[2]), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        acc += np.tensordot(src[:, xmin:xmin + cnt, :], kk[xx, :cnt].astype(np.int64), axes=([1], [0]))
        out[:, xx, :] = _clip8(acc)
    return out


def resize(img, out_w=224, out_h=224, filt=BILINEAR):
    """PIL.Image.resize((out_w,out_h), filt) on an (H, W, C) uint8 array.

    Pillow runs the horizontal pass first (uint8 intermediate), then the vertical pass,
    and skips a pass whose size is unchanged."""
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    if w != out_w:
        img = _resample_axis1(img, out_w, filt)
    if h != out_h:
        img = _resample_axis1(img.transpose(1, 0, 2), out_h, filt).transpose(1, 0, 2)
    return np.ascontiguousarray(img)
