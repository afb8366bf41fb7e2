"""Oracle (CPU, numpy) for A5/A6: dense Farneback optical flow and flow colouring.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

The reference calls a third-party dependency that is not under /root/reference:
opencv-python==4.9.0.80 (requirements.txt:73):
    cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
        src/main_fragment_layerstack.py:313-315
    flow_to_rgb: cv2.cartToPolar / cv2.normalize / cv2.cvtColor(HSV2BGR)
        src/main_fragment_layerstack.py:162-175
This file restates the published algorithm (OpenCV modules/video/src/optflowgf.cpp,
modules/imgproc smooth/resize/color_hsv) as spelled out in SURVEY.md 8(a) "Spec A5" and
row A6.  tests/test_oracle_flow.py pins it against the cv2 build in the image and the
reference's shipped *_residual_of.png fixtures.
"""
import numpy as np

F32 = np.float32

PYR_SCALE, LEVELS, WINSIZE, ITERATIONS, POLY_N, POLY_SIGMA = 0.5, 3, 15, 3, 5, 1.2
MIN_SIZE = 32
BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], dtype=F32)


def cv_round(x):
    """cvRound: round half to even."""
    return int(np.rint(x))


# --------------------------------------------------------------------------- primitives
def gaussian_taps(ksize, sigma):
    """cv::getGaussianKernel for float32 images; fixed table for sigma<=0, ksize=3."""
    if sigma <= 0:
        assert ksize == 3
        return np.array([0.25, 0.5, 0.25], dtype=F32)
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    k /= k.sum()
    return k.astype(F32)


def _reflect101(idx, n):
    idx = np.abs(idx)
    idx = np.where(idx >= n, 2 * (n - 1) - idx, idx)
    return idx


def gaussian_blur(img, ksize, sigma):
    """cv::GaussianBlur(f32, (k,k), sigma), BORDER_REFLECT_101, separable rows then cols."""
    taps = gaussian_taps(ksize, sigma)
    r = ksize // 2
    h, w = img.shape
    xs = _reflect101(np.arange(-r, w + r), w)
    padded = img[:, xs]
    tmp = np.zeros_like(img, dtype=F32)
    for k in range(ksize):
        tmp += taps[k] * padded[:, k:k + w]
    ys = _reflect101(np.arange(-r, h + r), h)
    padded = tmp[ys, :]
    out = np.zeros_like(img, dtype=F32)
    for k in range(ksize):
        out += taps[k] * padded[k:k + h, :]
    return out


def _linear_coords(n_in, n_out):
    scale = n_in / n_out
    src = (np.arange(n_out, dtype=np.float64) + 0.5) * scale - 0.5
    i0 = np.floor(src).astype(np.int64)
    a = (src - i0).astype(F32)
    lo = i0 < 0
    i0[lo] = 0
    a[lo] = 0
    hi = i0 >= n_in - 1
    i1 = i0 + 1
    i0[hi] = n_in - 1
    i1[hi] = n_in - 1
    a[hi] = 0
    return i0, i1, a


def resize_linear(img, w_out, h_out):
    """cv::resize(f32, INTER_LINEAR): half-pixel centres, no antialiasing; x then y."""
    h_in, w_in = img.shape[:2]
    if (w_in, h_in) == (w_out, h_out):
        return img.astype(F32, copy=True)
    x0, x1, ax = _linear_coords(w_in, w_out)
    y0, y1, ay = _linear_coords(h_in, h_out)
    if img.ndim == 3:
        ax = ax[None, :, None]
        ay = ay[:, None, None]
    else:
        ax = ax[None, :]
        ay = ay[:, None]
    rows = img[:, x0] * (F32(1) - ax) + img[:, x1] * ax
    out = rows[y0] * (F32(1) - ay) + rows[y1] * ay
    return out.astype(F32)


# ---------------------------------------------------------------- polynomial expansion
def prepare_gaussian(n=POLY_N, sigma=POLY_SIGMA):
    """FarnebackPrepareGaussian: taps g/xg/xxg (f32) and the four used entries of G^-1 (f64)."""
    x = np.arange(-n, n + 1)
    g = np.exp(-(x * x) / (2.0 * sigma * sigma)).astype(F32)
    s = 1.0 / float(g.astype(np.float64).sum())
    g = (g.astype(np.float64) * s).astype(F32)
    xg = (x * g.astype(np.float64)).astype(F32)
    xxg = (x * x * g.astype(np.float64)).astype(F32)
    G = np.zeros((6, 6))
    gd = g.astype(np.float64)
    for yy in range(-n, n + 1):
        for xx in range(-n, n + 1):
            ww = gd[yy + n] * gd[xx + n]
            G[0, 0] += ww
            G[1, 1] += ww * xx * xx
            G[3, 3] += ww * xx * xx * xx * xx
            G[5, 5] += ww * xx * xx * yy * yy
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return g, xg, xxg, inv[1, 1], inv[0, 3], inv[3, 3], inv[5, 5]


def poly_exp(img):
    """FarnebackPolyExp(I) -> R (h, w, 5) f32: [y, x, yy, xx, xy] coefficients."""
    n = POLY_N
    g, xg, xxg, ig11, ig03, ig33, ig55 = prepare_gaussian()
    h, w = img.shape
    img = img.astype(F32)
    # vertical pass, f32 accumulators, rows clamped
    r0 = img * g[n]
    r1 = np.zeros_like(img)
    r2 = np.zeros_like(img)
    ys = np.arange(h)
    for k in range(1, n + 1):
        up = img[np.maximum(ys - k, 0)]
        dn = img[np.minimum(ys + k, h - 1)]
        p = up + dn
        r0 = r0 + g[n + k] * p
        r1 = r1 + xg[n + k] * (dn - up)
        r2 = r2 + xxg[n + k] * p
    # horizontal pass, f64 accumulators, columns replicate-padded
    xs = np.clip(np.arange(-n, w + n), 0, w - 1)
    p0 = r0[:, xs].astype(np.float64)
    p1 = r1[:, xs].astype(np.float64)
    p2 = r2[:, xs].astype(np.float64)
    c = slice(n, n + w)
    gd, xgd, xxgd = g.astype(np.float64), xg.astype(np.float64), xxg.astype(np.float64)
    b1 = p0[:, c] * gd[n]
    b3 = p1[:, c] * gd[n]
    b5 = p2[:, c] * gd[n]
    b2 = np.zeros_like(b1)
    b4 = np.zeros_like(b1)
    b6 = np.zeros_like(b1)
    for k in range(1, n + 1):
        plus = slice(n + k, n + k + w)
        minus = slice(n - k, n - k + w)
        tg = p0[:, plus] + p0[:, minus]
        b1 += tg * gd[n + k]
        b4 += tg * xxgd[n + k]
        b2 += (p0[:, plus] - p0[:, minus]) * xgd[n + k]
        b3 += (p1[:, plus] + p1[:, minus]) * gd[n + k]
        b6 += (p1[:, plus] - p1[:, minus]) * xgd[n + k]
        b5 += (p2[:, plus] + p2[:, minus]) * gd[n + k]
    R = np.empty((h, w, 5), dtype=F32)
    R[..., 0] = (b3 * ig11).astype(F32)
    R[..., 1] = (b2 * ig11).astype(F32)
    R[..., 2] = (b1 * ig03 + b5 * ig33).astype(F32)
    R[..., 3] = (b1 * ig03 + b4 * ig33).astype(F32)
    R[..., 4] = (b6 * ig55).astype(F32)
    return R


def update_matrices(R0, R1, flow):
    """FarnebackUpdateMatrices -> M (h, w, 5) f32."""
    h, w = flow.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w]
    dx = flow[..., 0]
    dy = flow[..., 1]
    fx = (xs.astype(F32) + dx).astype(F32)
    fy = (ys.astype(F32) + dy).astype(F32)
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(F32)).astype(F32)
    fy = (fy - y1.astype(F32)).astype(F32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xc = np.clip(x1, 0, w - 2)
    yc = np.clip(y1, 0, h - 2)
    one = F32(1)
    a00 = (one - fx) * (one - fy)
    a01 = fx * (one - fy)
    a10 = (one - fx) * fy
    a11 = fx * fy
    p00 = R1[yc, xc]
    p01 = R1[yc, xc + 1]
    p10 = R1[yc + 1, xc]
    p11 = R1[yc + 1, xc + 1]
    interp = a00[..., None] * p00 + a01[..., None] * p01 + a10[..., None] * p10 + a11[..., None] * p11
    interp = interp.astype(F32)
    r2 = np.where(inside, interp[..., 0], F32(0)).astype(F32)
    r3 = np.where(inside, interp[..., 1], F32(0)).astype(F32)
    r4 = np.where(inside, (R0[..., 2] + interp[..., 2]) * F32(0.5), R0[..., 2]).astype(F32)
    r5 = np.where(inside, (R0[..., 3] + interp[..., 3]) * F32(0.5), R0[..., 3]).astype(F32)
    r6 = np.where(inside, (R0[..., 4] + interp[..., 4]) * F32(0.25), R0[..., 4] * F32(0.5)).astype(F32)
    r2 = (R0[..., 0] - r2) * F32(0.5)
    r3 = (R0[..., 1] - r3) * F32(0.5)
    r2 = (r2 + (r4 * dy + r6 * dx)).astype(F32)
    r3 = (r3 + (r6 * dy + r5 * dx)).astype(F32)
    sx = np.ones(w, dtype=F32)
    sy = np.ones(h, dtype=F32)
    nb = len(BORDER)
    for i in range(min(nb, w)):
        sx[i] *= BORDER[i]
        sx[w - 1 - i] *= BORDER[i]
    for i in range(min(nb, h)):
        sy[i] *= BORDER[i]
        sy[h - 1 - i] *= BORDER[i]
    scale = (sy[:, None] * sx[None, :]).astype(F32)
    r2, r3, r4, r5, r6 = (v * scale for v in (r2, r3, r4, r5, r6))
    M = np.empty((h, w, 5), dtype=F32)
    M[..., 0] = r4 * r4 + r6 * r6
    M[..., 1] = (r4 + r5) * r6
    M[..., 2] = r5 * r5 + r6 * r6
    M[..., 3] = r4 * r2 + r6 * r3
    M[..., 4] = r6 * r2 + r5 * r3
    return M


def box_solve(M, block=WINSIZE):
    """FarnebackUpdateFlow_Blur: 15x15 box sum (f64, replicate border) + 2x2 solve."""
    h, w = M.shape[:2]
    m = block // 2
    Md = M.astype(np.float64)
    ys = np.clip(np.arange(-m, h + m), 0, h - 1)
    xs = np.clip(np.arange(-m, w + m), 0, w - 1)
    P = Md[ys][:, xs]
    c = np.cumsum(P, axis=0)
    c = np.concatenate([np.zeros((1,) + c.shape[1:]), c], axis=0)
    v = c[block:] - c[:-block]
    c = np.cumsum(v, axis=1)
    c = np.concatenate([np.zeros((c.shape[0], 1, 5)), c], axis=1)
    S = (c[:, block:] - c[:, :-block]) * (1.0 / (block * block))
    g11, g12, g22, h1, h2 = (S[..., i] for i in range(5))
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    flow = np.empty((h, w, 2), dtype=F32)
    flow[..., 0] = ((g11 * h2 - g12 * h1) * idet).astype(F32)
    flow[..., 1] = ((g22 * h1 - g12 * h2) * idet).astype(F32)
    return flow


def pyramid_plan(h, w):
    """Levels actually used and per-level (scale, sigma, ksize, w_k, h_k), coarse -> fine."""
    levels = 0
    scale = 1.0
    while levels < LEVELS:
        scale *= PYR_SCALE
        if w * scale < MIN_SIZE or h * scale < MIN_SIZE:
            break
        levels += 1
    plan = []
    for k in range(levels, -1, -1):
        scale = PYR_SCALE ** k
        sigma = (1.0 / scale - 1.0) * 0.5
        ksize = max(cv_round(sigma * 5) | 1, 3)
        plan.append((scale, sigma, ksize, cv_round(w * scale), cv_round(h * scale)))
    return plan


def farneback(gray0, gray1):
    """cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0) -> (H, W, 2) f32."""
    h, w = gray0.shape
    imgs = [gray0.astype(F32), gray1.astype(F32)]
    flow = None
    for (scale, sigma, ksize, wk, hk) in pyramid_plan(h, w):
        if flow is None:
            flow = np.zeros((hk, wk, 2), dtype=F32)
        else:
            flow = resize_linear(flow, wk, hk) * F32(1.0 / PYR_SCALE)
        R = []
        for im in imgs:
            blurred = gaussian_blur(im, ksize, sigma)
            R.append(poly_exp(resize_linear(blurred, wk, hk)))
        M = update_matrices(R[0], R[1], flow)
        for it in range(ITERATIONS):
            flow = box_solve(M)
            if it < ITERATIONS - 1:
                M = update_matrices(R[0], R[1], flow)
    return flow


# ------------------------------------------------------------------------ flow colouring
def fast_atan2_deg(y, x):
    """cv::fastAtan2 polynomial (degrees, [0,360)), as used by cartToPolar (f32)."""
    p1, p3, p5, p7 = F32(0.9997878412794807 * (180 / np.pi)), F32(-0.3258083974640975 * (180 / np.pi)), \
        F32(0.1555786518463281 * (180 / np.pi)), F32(-0.04432655554792128 * (180 / np.pi))
    ax, ay = np.abs(x).astype(F32), np.abs(y).astype(F32)
    eps = F32(2.220446049250313e-16)
    swap = ax < ay
    num = np.where(swap, ax, ay)
    den = np.where(swap, ay, ax) + eps
    c = (num / den).astype(F32)
    c2 = (c * c).astype(F32)
    a = ((((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c).astype(F32)
    a = np.where(swap, F32(90) - a, a)
    a = np.where(x < 0, F32(180) - a, a)
    a = np.where(y < 0, F32(360) - a, a)
    return a.astype(F32)


def hsv2bgr_u8(hsv):
    """cv2.cvtColor(uint8 HSV -> BGR), H in [0,180): float formula then x255 and truncate
    (matches cv2's SIMD path on all 181x256 (H,V) inputs with S=255)."""
    hh = hsv[..., 0].astype(F32) * F32(6.0 / 180.0)
    s = hsv[..., 1].astype(F32) * F32(1.0 / 255.0)
    v = hsv[..., 2].astype(F32) * F32(1.0 / 255.0)
    sector = np.floor(hh).astype(np.int32)
    f = (hh - sector.astype(F32)).astype(F32)
    sector = sector % 6
    one = F32(1)
    t0 = v
    t1 = (v * (one - s)).astype(F32)
    t2 = (v * (one - s * f)).astype(F32)
    t3 = (v * (one - s * (one - f))).astype(F32)
    tab = np.stack([t0, t1, t2, t3], axis=-1)
    sector_to = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    idx = sector_to[sector]
    b = np.take_along_axis(tab, idx[..., 0:1], -1)[..., 0]
    g = np.take_along_axis(tab, idx[..., 1:2], -1)[..., 0]
    r = np.take_along_axis(tab, idx[..., 2:3], -1)[..., 0]
    out = np.stack([b, g, r], axis=-1) * F32(255)
    return np.clip(out, 0, 255).astype(np.uint8)


def _cv_normalize_minmax_255(m):
    """cv2.normalize(m, None, 0, 255, NORM_MINMAX) on float32 (third-party, opencv-python 4.9.0.80 pinned by the reference's
    requirements.txt:73; behaviour probed on cv2 4.13): scale = (float)(255 * (1 / (max - min))) with the division in double,
    shift = -(min * scale) with the float32 scale, then dst = fma(src, scale, shift) in float32 (convertTo's 32f -> 32f
    path).  Bit-identical to cv2 on every probe (tests/test_oracle_flow.py::test_normalize_twice_is_cv2_exact)."""
    mn, mx = float(m.min()), float(m.max())
    sc = F32(255.0 * (1.0 / (mx - mn)) if mx - mn > 2.220446049250313e-16 else 0.0)
    a, b = np.float64(sc), np.float64(F32(-(F32(mn) * sc)))
    return (m.astype(np.float64) * a + b).astype(F32)       # the double product of two floats is exact -> one rounding = fma


def cv_magnitude(dx, dy):
    """cv2.cartToPolar's magnitude: sqrt(fma(x, x, fl(y*y))) in float32 (bit-identical to cv2 4.13 on every probe)."""
    yy = (dy.astype(F32) * dy.astype(F32)).astype(F32)
    return np.sqrt((dx.astype(np.float64) * dx.astype(np.float64) + yy.astype(np.float64)).astype(F32)).astype(F32)


def flow_to_rgb(flow):
    """flow_to_rgb - src/main_fragment_layerstack.py:162-175 (returns BGR like the reference).  The reference
    normalises the magnitude TWICE (mag, then again for hsv[..., 2]); both passes are restated in cv2's float32
    arithmetic."""
    dx = flow[..., 0].astype(F32)
    dy = flow[..., 1].astype(F32)
    mag = cv_magnitude(dx, dy)
    ang = (fast_atan2_deg(dy, dx) * F32(np.pi / 180.0)).astype(F32)
    magn = _cv_normalize_minmax_255(_cv_normalize_minmax_255(mag))
    hue = ang * 180 / np.pi / 2            # float32 array arithmetic, as in the reference
    hsv = np.zeros(flow.shape[:2] + (3,), dtype=np.uint8)
    hsv[..., 0] = hue.astype(np.uint8) if hue.dtype != np.uint8 else hue
    hsv[..., 1] = 255
    hsv[..., 2] = np.clip(magn, 0, 255).astype(np.uint8)
    return hsv2bgr_u8(hsv)
