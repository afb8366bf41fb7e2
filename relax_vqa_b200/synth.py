"""Seeded synthetic clips shaped like the reference's datasets.

A clip is T sampled pairs of BGR uint8 frames: ``frames[t]`` is sampled frame t*k and
``nexts[t]`` its successor t*k+1 (k = int(fps/2), src/main_fragment_layerstack.py:274-277).
Texture = Gaussian-blurred uniform noise; the successor is the same texture under a global
sub-pixel shift, with one block moved by (5,10) px and +-3 noise, so residuals, patch sums
and optical flow are all non-trivial.
"""
import numpy as np

CONFIGS = {
    # name: (H, W, n_frames, fps)  -> pairs = number of sampled pairs
    "540p-8s": (540, 960, 240, 29.97),
    "1080p-10s": (1080, 1920, 300, 29.97),
    "2160p-20s": (2160, 3840, 600, 29.97),
}


def sampled_counts(n_frames, fps):
    """(#sampled frames, #pairs) the reference's two ffmpeg select filters produce
    (src/video_frames_extract.py:6-27,51-74)."""
    k = int(np.ceil(fps / 2)) if fps < 2 else int(fps / 2)
    n_full = -(-n_frames // k)
    n_next = -(-(n_frames - 1) // k)
    return n_full, min(n_full, n_next)


def make_pair(rng, H, W):
    import cv2
    pad = 16
    noise = rng.uniform(0, 255, (H + 2 * pad, W + 2 * pad, 3)).astype(np.float32)
    base = cv2.GaussianBlur(noise, (0, 0), float(rng.uniform(3.0, 4.0)))
    base = (base - base.mean()) * 6.0 + 128.0
    dx, dy = rng.uniform(-3, 3, 2)
    Mshift = np.float32([[1, 0, dx], [0, 1, dy]])
    moved = cv2.warpAffine(base, Mshift, (base.shape[1], base.shape[0]), flags=cv2.INTER_LINEAR,
                           borderMode=cv2.BORDER_REFLECT_101)
    bh, bw = min(130, H // 3), min(160, W // 3)
    by, bx = int(rng.integers(pad, H - bh - 12)), int(rng.integers(pad, W - bw - 12))
    moved[by + 5:by + 5 + bh, bx + 10:bx + 10 + bw] = base[by:by + bh, bx:bx + bw]
    moved = moved + rng.uniform(-3, 3, moved.shape).astype(np.float32)
    f0 = np.clip(base[pad:pad + H, pad:pad + W], 0, 255).astype(np.uint8)
    f1 = np.clip(moved[pad:pad + H, pad:pad + W], 0, 255).astype(np.uint8)
    return f0, f1


def make_clip(seed, H, W, pairs):
    """-> (frames, nexts): uint8 arrays (pairs, H, W, 3), BGR."""
    rng = np.random.default_rng(seed)
    fr, nx = zip(*(make_pair(rng, H, W) for _ in range(pairs)))
    return np.stack(fr), np.stack(nx)
