"""Pins oracle/sampler.py against the real libswscale: the OpenCV wheel in this image bundles
FFmpeg, and cv2.VideoCapture on a raw .yuv file (rawvideo demuxer, options passed through
OPENCV_FFMPEG_CAPTURE_OPTIONS) runs swscale's unscaled yuv420p -> bgr24 converter - the conversion ffmpeg applies before
the reference's PNGs are written.  Runs in a subprocess because the option string is read when the backend starts."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import sampler as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r'''
import json, os, sys
import numpy as np
W, H, n, path = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
os.environ["OPENCV_FFMPEG_CAPTURE_OPTIONS"] = f"video_size;{W}x{H}|pixel_format;yuv420p|framerate;25"
import cv2
sys.path.insert(0, sys.argv[5])
from oracle import sampler as S
cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
fb = H * W * 3 // 2
raw = np.fromfile(path, dtype=np.uint8)
bad, worst, got = 0, 0, 0
while True:
    ok, f = cap.read()
    if not ok:
        break
    ref = S.yuv420p_to_bgr(*S.split_planes(raw[got * fb:(got + 1) * fb], H, W))
    d = np.abs(ref.astype(int) - f.astype(int))
    bad += int((d != 0).sum()); worst = max(worst, int(d.max())); got += 1
print(json.dumps(dict(opened=cap.isOpened() or got > 0, frames=got, mismatching=bad, worst=worst)))
'''


@pytest.mark.parametrize("wh", [(64, 48), (1920, 1080), (100, 38)])
def test_yuv420p_to_bgr_is_swscale_exact(tmp_path, wh):
    W, H = wh
    rng = np.random.default_rng(W)
    n = 3
    frames = rng.integers(0, 256, (n, H * W * 3 // 2), dtype=np.uint8)           # full-range garbage: every table entry / clamp
    frames[1, :H * W] = np.clip(frames[1, :H * W], 16, 235)                       # and one legal-range luma plane
    path = str(tmp_path / "a.yuv")
    frames.tofile(path)
    out = subprocess.run([sys.executable, "-c", PROBE, str(W), str(H), str(n), path, ROOT], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-800:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    if not r["opened"] or r["frames"] == 0:
        pytest.skip("this OpenCV build cannot open raw yuv through FFmpeg")
    assert r["frames"] == n and r["mismatching"] == 0 and r["worst"] == 0, r


def test_frame_selection_matches_the_select_filters():
    full, nxt = S.selected_indices(300, 14)                 # 1080p-10s: 300 frames @29.97
    assert full[:3] == [0, 14, 28] and nxt[:3] == [1, 15, 29] and len(full) == 22 and len(nxt) == 22
    full, nxt = S.selected_indices(15, 14)                  # (N-1) % k == 0: one pair fewer than sampled frames
    assert full == [0, 14] and nxt == [1]
    from relax_vqa_b200 import synth
    for nf, fps in ((240, 29.97), (300, 29.97), (600, 29.97), (15, 29.97), (1, 30.0)):
        k = int(fps / 2)
        full, nxt = S.selected_indices(nf, k)
        assert (len(full), min(len(full), len(nxt))) == synth.sampled_counts(nf, fps)


def test_sample_yuv420p_oracle_layout(tmp_path):
    W, H, n, k = 32, 16, 7, 3
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (n, H * W * 3 // 2), dtype=np.uint8)
    path = str(tmp_path / "v.yuv")
    frames.tofile(path)
    fr, nx = S.sample_yuv420p(path, W, H, k)
    assert fr.shape == (3, H, W, 3) and nx.shape == (2, H, W, 3)           # frames 0,3,6 and 1,4
    assert np.array_equal(nx[1], S.yuv420p_to_bgr(*S.split_planes(frames[4], H, W)))
