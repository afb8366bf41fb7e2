"""Pins oracle/farneback.py (A5, A6) against cv2 (the reference's dependency) and the
shipped *_residual_of.png fixtures.  Float tolerances are stated per assertion."""
import os

import cv2
import numpy as np
import pytest

from oracle import farneback as FB
from oracle import fragments as F
from relax_vqa_b200 import synth


def _cv_flow(g0, g1):
    return cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)


def _cv_flow_to_rgb(flow):   # body of src/main_fragment_layerstack.py:162-175
    mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
    mag = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
    hue = ang * 180 / np.pi / 2
    hsv = np.zeros((flow.shape[0], flow.shape[1], 3), dtype=np.uint8)
    hsv[..., 0] = hue
    hsv[..., 1] = 255
    hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
    return cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)


@pytest.mark.parametrize("hw", [(120, 160), (272, 480), (135, 241)])
def test_farneback_vs_cv2(hw):
    fr, nx = synth.make_clip(3, hw[0], hw[1], 1)
    g0, g1 = F.bgr2gray(fr[0]), F.bgr2gray(nx[0])
    ref = _cv_flow(g0, g1)
    got = FB.farneback(g0, g1)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.abs(got - ref).max() < 5e-4          # px; observed <= 1e-4
    assert np.abs(got - ref).mean() < 5e-6


def test_pyramid_plan():
    assert [p[2] for p in FB.pyramid_plan(540, 960)] == [19, 9, 3, 3]
    assert [(p[3], p[4]) for p in FB.pyramid_plan(540, 960)] == [(120, 68), (240, 135), (480, 270), (960, 540)]
    assert len(FB.pyramid_plan(100, 150)) == 2      # 25x37 < 32 stops the pyramid early


@pytest.mark.parametrize("hw,seed", [((272, 480), 5), ((544, 960), 6), ((128, 256), 7)])
def test_flow_to_rgb_vs_cv2(hw, seed):
    """Widths that are multiples of cv2's SIMD width: the whole colouring (magnitude, both normalize passes, hue,
    HSV2BGR) is bit-exact up to the 1-ulp differences of cv2's vectorised fastAtan2 (< 1e-5 of the pixels)."""
    fr, nx = synth.make_clip(seed, hw[0], hw[1], 1)
    flow = _cv_flow(F.bgr2gray(fr[0]), F.bgr2gray(nx[0]))
    got, ref = FB.flow_to_rgb(flow), _cv_flow_to_rgb(flow)
    bad = (got != ref).any(-1)
    assert bad.mean() < 1e-5, bad.sum()


def test_flow_to_rgb_vs_cv2_ragged_width():
    # cv2's scalar row-tail path rounds instead of truncating: +-1 on the last < 32 px of a row
    rng = np.random.default_rng(0)
    flow = (rng.standard_normal((123, 211, 2)) * 2).astype(np.float32)
    got, ref = FB.flow_to_rgb(flow), _cv_flow_to_rgb(flow)
    bad = (got != ref).any(-1)
    assert not bad[:, :211 - 32].any() and np.abs(got.astype(int) - ref.astype(int)).max() <= 1


@pytest.mark.parametrize("seed", range(12))
def test_normalize_twice_is_cv2_exact(seed):
    """The reference normalises the magnitude twice (src/main_fragment_layerstack.py:164,172); the float32 restatement
    of cv2.normalize and of cartToPolar's magnitude is bit-identical to cv2 for both passes."""
    rng = np.random.default_rng(seed)
    flow = (rng.standard_normal((97, 161, 2)) * rng.uniform(0.05, 8)).astype(np.float32)
    mag, _ = cv2.cartToPolar(flow[..., 0], flow[..., 1])
    assert np.array_equal(FB.cv_magnitude(flow[..., 0], flow[..., 1]), mag)
    m1 = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
    m2 = cv2.normalize(m1, None, 0, 255, cv2.NORM_MINMAX)
    assert np.array_equal(FB._cv_normalize_minmax_255(mag), m1)
    assert np.array_equal(FB._cv_normalize_minmax_255(m1), m2)


@pytest.mark.parametrize("idx", [2, 3])
def test_shipped_flow_fixture(example_dir, idx):
    d = os.path.join(example_dir, "original_5636101558")
    a = cv2.imread(os.path.join(d, f"5636101558_{idx}.png"))
    b = cv2.imread(os.path.join(d, f"5636101558_{idx}_next.png"))
    stored = cv2.imread(os.path.join(d, f"5636101558_{idx}_residual_of.png"))
    got = FB.flow_to_rgb(FB.farneback(F.bgr2gray(a), F.bgr2gray(b)))
    diff = np.abs(got.astype(int) - stored.astype(int))
    # authors' OpenCV 4.9 vs this restatement: a handful of +-1 pixels
    assert (diff != 0).any(-1).sum() <= 16 and diff.max() <= 2
