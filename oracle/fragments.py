"""Oracle (CPU, numpy) for the integer stages A1-A4, A7, A8 of SURVEY.md section 8(a).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Each function restates the reference's algorithm; none imports cv2 or the reference.
Reference files are cited relative to /root/reference/.
"""
import numpy as np

PATCH = 16
TARGET = 224
TOP_N = 196


def absdiff(a, b):
    """cv2.absdiff(img_next, img_original) - src/main_fragment_layerstack.py:302."""
    a = np.asarray(a)
    b = np.asarray(b)
    return np.abs(a.astype(np.int16) - b.astype(np.int16)).astype(np.uint8)


def bgr2gray(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) on uint8 - src/main_fragment_layerstack.py:313-314.

    OpenCV's 8-bit path is fixed point: (B*3735 + G*19235 + R*9798 + 2^14) >> 15
    (bit-exact vs cv2 4.13)."""
    img = np.asarray(img).astype(np.uint32)
    g = (img[..., 0] * 3735 + img[..., 1] * 19235 + img[..., 2] * 9798 + (1 << 14)) >> 15
    return g.astype(np.uint8)


def patch_sums(residual, patch=PATCH):
    """get_patch_diff - src/main_fragment_layerstack.py:177-189.

    Crops to a multiple of `patch`, sums |residual| (a no-op on uint8) over each
    patch and all channels.  Returned as float64 (gh, gw) like the reference; every
    value is an exact integer <= 16*16*3*255."""
    h, w = residual.shape[:2]
    gh, gw = h // patch, w // patch
    r = residual[:gh * patch, :gw * patch].astype(np.uint64)
    r = r.reshape(gh, patch, gw, patch, -1)
    return r.sum(axis=(1, 3, 4)).astype(np.float64)


def topk_positions(diff, top_n=TOP_N):
    """Selection half of extract_important_patches - src/main_fragment_layerstack.py:193-195.

    The reference uses np.argsort(-diff.ravel()) (unstable).  The build's tie contract
     is value-descending, flat-index-ascending, i.e. the stable
    argsort; the result is then sorted into raster order.  Returns list[(y, x)]."""
    flat = np.asarray(diff).ravel()
    order = np.argsort(-flat, kind="stable")[:top_n]
    gw = diff.shape[1]
    return sorted((int(i // gw), int(i % gw)) for i in order)


def cut_is_tied(diff, top_n=TOP_N):
    """True when the top_n / top_n+1 boundary falls inside a run of equal sums, i.e.
    when the reference's own (unstable) choice is not a portable fact."""
    flat = np.sort(np.asarray(diff).ravel())[::-1]
    return flat.size > top_n and flat[top_n - 1] == flat[top_n]


def gather_fragment(frame, positions, patch=PATCH, target=TARGET):
    """Patch copy of extract_important_patches (:197-210) and
    get_original_frame_patches (:212-230): patch j -> canvas cell (j // 14, j % 14)."""
    out = np.zeros((target, target, frame.shape[2]), dtype=frame.dtype)
    per_row = target // patch
    for j, (y, x) in enumerate(positions):
        ty, tx = (j // per_row) * patch, (j % per_row) * patch
        out[ty:ty + patch, tx:tx + patch] = frame[y * patch:(y + 1) * patch, x * patch:(x + 1) * patch]
    return out


def extract_important_patches(residual, diff, patch=PATCH, target=TARGET, top_n=TOP_N):
    """extract_important_patches - src/main_fragment_layerstack.py:191-210."""
    pos = topk_positions(diff, top_n)
    return gather_fragment(residual, pos, patch, target), pos


def process_patches(residual, patch=PATCH, target=TARGET, top_n=TOP_N):
    """process_patches (:232-240) without the path bookkeeping: (fragment, positions, sums)."""
    d = patch_sums(residual, patch)
    frag, pos = extract_important_patches(residual, d, patch, target, top_n)
    return frag, pos, d


def merge_fragments(a, b):
    """merge_fragments - src/main_fragment_layerstack.py:242-245.

    cv2.addWeighted(a, .5, b, .5, 0) on uint8 = (a+b)/2 rounded half-to-even
    (1,2->2; 2,3->2; 3,4->4; 255,254->254)."""
    s = a.astype(np.uint16) + b.astype(np.uint16)
    half = s >> 1
    odd = s & 1
    return (half + (odd & (half & 1))).astype(np.uint8)
