"""Oracle (CPU) for frame sampling: which frames the reference's two ffmpeg select filters keep and
how libswscale turns a raw yuv420p frame into the BGR pixels cv2.imread later returns.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

The arithmetic lives in a third-party dependency that is not under /root/reference: FFmpeg (the reference shells out to an
`ffmpeg` binary, src/video_frames_extract.py:6-100; no version pinned).  Restated from libswscale's unscaled yuv420p -> rgb24 /
bgr24 converter, x86 SIMD path (libswscale/x86/yuv_2_rgb.asm, coefficients from ff_yuv2rgb_c_init_tables in
libswscale/yuv2rgb.c, ITU-R BT.601 limited range = SWS_CS_DEFAULT).  PINNED against the real thing: the OpenCV wheel in this
image bundles FFmpeg 8 (avcodec 62 / swscale 9.1), and cv2.VideoCapture on a raw .yuv file runs exactly this converter
(tests/test_oracle_sampler.py: bit-identical on random planes).  Not covered: odd heights (swscale then uses its C table
path) and the other pixel formats of the `pixfmt` argument.
"""
import numpy as np


def _round_to_int16(f):
    return int(np.clip((f + (1 << 15)) >> 16, -32768, 32767))


# ff_yuv2rgb_c_init_tables: 16.16 coefficients of SWS_CS_ITU601, limited range; then x 2^13 rounded to int16
_CY = (65536 * 255) // 219
Y_COEFF, VR_COEFF, UB_COEFF, VG_COEFF, UG_COEFF = (_round_to_int16(c * 8192) for c in (_CY, 104597, 132201, -53279, -25675))
Y_OFFSET = _round_to_int16((16 << 16) * 8)          # 128 = 16 << 3
assert (Y_COEFF, VR_COEFF, UB_COEFF, VG_COEFF, UG_COEFF, Y_OFFSET) == (9539, 13075, 16525, -6660, -3209, 128)


def yuv420p_to_bgr(Y, U, V):
    """Y (H,W), U, V (H/2,W/2) uint8 -> (H,W,3) uint8 BGR.  pmulhw = arithmetic >> 16 of the signed product."""
    Yl = Y.astype(np.int64)
    Un = np.repeat(np.repeat(U, 2, 0), 2, 1).astype(np.int64)      # chroma is not interpolated
    Vn = np.repeat(np.repeat(V, 2, 0), 2, 1).astype(np.int64)
    yp = (((Yl << 3) - Y_OFFSET) * Y_COEFF) >> 16
    up, vp = (Un << 3) - 0x400, (Vn << 3) - 0x400
    r = yp + ((vp * VR_COEFF) >> 16)
    g = yp + ((up * UG_COEFF) >> 16) + ((vp * VG_COEFF) >> 16)
    b = yp + ((up * UB_COEFF) >> 16)
    return np.clip(np.stack([b, g, r], axis=-1), 0, 255).astype(np.uint8)


def split_planes(frame_bytes, H, W):
    a = np.frombuffer(frame_bytes, dtype=np.uint8) if not isinstance(frame_bytes, np.ndarray) else frame_bytes
    n = H * W
    return a[:n].reshape(H, W), a[n:n + n // 4].reshape(H // 2, W // 2), a[n + n // 4:n + n // 2].reshape(H // 2, W // 2)


def selected_indices(n_frames, frame_interval):
    """(frames kept by select='not(mod(n,k))', frames kept by select='not(mod(n-1,k))') - 0-based decode order; the
    reference numbers the PNGs from 1 in this order (src/video_frames_extract.py:9,61)."""
    k = int(frame_interval)
    return [n for n in range(n_frames) if n % k == 0], [n for n in range(n_frames) if (n - 1) % k == 0]


def sample_yuv420p(path, W, H, frame_interval):
    """-> (frames [Tf,H,W,3], nexts [Tp,H,W,3]) BGR uint8: what cv2.imread returns for the reference's sampled PNGs."""
    fb = H * W * 3 // 2
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    n = raw.size // fb
    full, nxt = selected_indices(n, frame_interval)
    conv = lambda i: yuv420p_to_bgr(*split_planes(np.asarray(raw[i * fb:(i + 1) * fb]), H, W))
    pairs = min(len(full), len(nxt))
    z = np.zeros((0, H, W, 3), np.uint8)
    return (np.stack([conv(i) for i in full]) if full else z), (np.stack([conv(i) for i in nxt[:pairs]]) if pairs else z)
