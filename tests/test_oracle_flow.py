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


def test_flow_to_rgb_vs_cv2():
    fr, nx = synth.make_clip(5, 272, 480, 1)
    flow = _cv_flow(F.bgr2gray(fr[0]), F.bgr2gray(nx[0]))
    got, ref = FB.flow_to_rgb(flow), _cv_flow_to_rgb(flow)
    bad = (got != ref).any(-1)
    # cv2's scalar row-tail path rounds instead of truncating: allow +-1 on a few px
    assert bad.mean() < 1e-3 and np.abs(got.astype(int) - ref.astype(int)).max() <= 1


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
