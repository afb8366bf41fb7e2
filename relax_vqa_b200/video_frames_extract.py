"""Drop-in for src/video_frames_extract.py (= src/extractor/vf_extract.py): frame sampling feeding the hot path
(SURVEY.md 8(f) row 1).

The reference shells out to ffmpeg twice per video: ``select='not(mod(n,k))'`` -> ``{vid}_{i}.png`` and
``select='not(mod(n-1,k))'`` -> ``{vid}_{i}_next.png`` (k = int(fps / 2)), then reads the PNGs back with cv2.  Here:

* raw yuv420p input (the reference's ``live_qualcomm`` branch, :29-49 / :76-100): the file is memory-mapped, only the
  selected frames are staged through pinned memory and converted on the GPU by ``b200vqa_yuv420p_to_bgr`` - bit-identical
  to what ffmpeg's libswscale writes into the PNGs - straight into an ``engine.Clip`` (no PNG round trip);
* container formats (.mp4 / .mkv ..., :6-27 / :51-74): decoded on the host by the FFmpeg libraries inside OpenCV
  (cv2.VideoCapture) - the same decoder family as the reference's ffmpeg binary; there is no NVDEC / ffmpeg binary in this
  image, and decoding is upstream of the hot path.  Frames are selected by index in decode order, like the select filter.
  (ffmpeg converts the decoded frame to rgb24 with its default bicubic chroma upsampling, OpenCV asks swscale for bgr24
  with SWS_BICUBIC: the same scaler; parity with an ffmpeg-written PNG is unpinned here for lack of an ffmpeg binary.)

``process_video`` / ``process_video_residual`` keep the reference's signatures and write the same PNG files, so its
PNG-folder workflow still works; ``sample_clip`` is the fast path used by ``demo_test.evaluate_video_quality``."""
import math
import os

import cv2
import numpy as np
import torch

from . import ops
from .engine import Clip


def frame_interval_of(framerate):
    """src/main_fragment_layerstack.py:274-277 (demo_test.py:76 uses int(framerate / 2) directly)."""
    return math.ceil(framerate / 2) if framerate < 2 else int(framerate / 2)


def selected_indices(n_frames, frame_interval):
    """0-based indices (decode order) kept by the two select filters (:16, :66)."""
    k = max(1, int(frame_interval))
    return [n for n in range(n_frames) if n % k == 0], [n for n in range(n_frames) if (n - 1) % k == 0]


def _device(device):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        from ._lib import B200VQAError
        raise B200VQAError("no CUDA device: relax_vqa_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def sample_yuv420p(video_path, video_width, video_height, frame_interval, device=None, pixfmt="yuv420p"):
    """Raw planar 4:2:0 file -> Clip(frames [Tf,H,W,3], nexts [Tp,H,W,3]) BGR uint8 on the device."""
    if pixfmt != "yuv420p":
        raise ValueError(f"raw sampling implements pix_fmt yuv420p (the reference's LIVE-Qualcomm input), not {pixfmt}")
    W, H = int(video_width), int(video_height)
    if W % 4 or H % 2:
        raise ValueError("yuv420p sampling needs W % 4 == 0 and an even H (swscale's unscaled SIMD converter)")
    fb = H * W * 3 // 2
    raw = np.memmap(video_path, dtype=np.uint8, mode="r")
    n = raw.size // fb
    if n == 0:
        raise ValueError(f"{video_path}: shorter than one {W}x{H} yuv420p frame")
    full, nxt = selected_indices(n, frame_interval)
    nxt = nxt[:min(len(full), len(nxt))]
    dev = _device(device)
    idx = full + nxt
    stage = torch.empty((len(idx), fb), dtype=torch.uint8, pin_memory=True)
    view = stage.numpy()
    for j, i in enumerate(idx):                       # only the sampled frames leave the page cache
        view[j] = raw[i * fb:(i + 1) * fb]
    bgr = ops.yuv420p_to_bgr(stage.to(dev, non_blocking=True), H, W)
    return Clip(bgr[:len(full)], bgr[len(full):])


def sample_video(video_path, frame_interval, device=None):
    """Container file -> Clip, decoded on the host by OpenCV's FFmpeg backend (frames in decode order)."""
    cap = cv2.VideoCapture(video_path, cv2.CAP_FFMPEG)
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open {video_path}")
    k = max(1, int(frame_interval))
    full, nxt, n = [], [], 0
    while True:
        ok = cap.grab()
        if not ok:
            break
        if n % k == 0 or (n - 1) % k == 0:
            ok, frame = cap.retrieve()
            if not ok:
                break
            (full if n % k == 0 else nxt).append(frame)
            if k == 1 and n >= 1:                     # every frame is both a sample and a successor
                nxt.append(frame)
        n += 1
    cap.release()
    if not full:
        raise ValueError(f"{video_path}: no frames decoded")
    nxt = nxt[:min(len(full), len(nxt))]
    dev = _device(device)
    f = torch.from_numpy(np.stack(full)).to(dev)
    nx = torch.from_numpy(np.stack(nxt)).to(dev) if nxt else torch.empty((0,) + tuple(f.shape[1:]), dtype=torch.uint8, device=dev)
    return Clip(f, nx)


def sample_clip(video_type, video_path, frame_interval, video_width=None, video_height=None, pixfmt="yuv420p", device=None):
    """The reference's dispatch (:103-121): raw yuv for live_qualcomm, the demuxer / decoder otherwise."""
    if video_type == "live_qualcomm" or str(video_path).lower().endswith(".yuv"):
        return sample_yuv420p(video_path, video_width, video_height, frame_interval, device, pixfmt)
    return sample_video(video_path, frame_interval, device)


# ------------------------------------------------------------------ reference-named entry points (PNG outputs)
def _write_pngs(clip, out_dir, video_name, want_full=True, want_next=False):
    os.makedirs(out_dir, exist_ok=True)
    if want_full:
        for i, img in enumerate(clip.frames.cpu().numpy()):
            cv2.imwrite(os.path.join(out_dir, f"{video_name}_{i + 1}.png"), img)
    if want_next:
        for i, img in enumerate(clip.nexts.cpu().numpy()):
            cv2.imwrite(os.path.join(out_dir, f"{video_name}_{i + 1}_next.png"), img)


def extract_frames_general(input_video_path, output_frame_directory, frame_interval):
    """ref :6-27."""
    name = os.path.splitext(os.path.basename(input_video_path))[0]
    _write_pngs(sample_video(input_video_path, frame_interval), output_frame_directory, name)


def extract_frames_yuv(input_video_path, output_frame_directory, frame_interval, video_width, video_height, pixfmt, framerate):
    """ref :29-49."""
    name = os.path.splitext(os.path.basename(input_video_path))[0]
    _write_pngs(sample_yuv420p(input_video_path, video_width, video_height, frame_interval, pixfmt=pixfmt), output_frame_directory, name)


def extract_frames_residual(video_path, sampled_path, frame_interval):
    """ref :51-74."""
    name = os.path.splitext(os.path.basename(video_path))[0]
    _write_pngs(sample_video(video_path, frame_interval), sampled_path, name, want_next=True)


def extract_frames_residual_yuv(video_path, sampled_path, frame_interval, video_width, video_height, pixfmt, framerate):
    """ref :76-100."""
    name = os.path.splitext(os.path.basename(video_path))[0]
    _write_pngs(sample_yuv420p(video_path, video_width, video_height, frame_interval, pixfmt=pixfmt), sampled_path, name, want_next=True)


def process_video(video_type, video_name, frame_interval, video_path, sampled_path, video_width, video_height, pixfmt, framerate):
    """ref :103-111."""
    os.makedirs(sampled_path, exist_ok=True)
    if video_type == 'live_qualcomm':
        extract_frames_yuv(video_path, sampled_path, frame_interval, video_width, video_height, pixfmt, framerate)
    else:
        extract_frames_general(video_path, sampled_path, frame_interval)


def process_video_residual(video_type, video_name, frame_interval, video_path, sampled_path, video_width, video_height, pixfmt, framerate):
    """ref :113-121."""
    os.makedirs(sampled_path, exist_ok=True)
    if video_type == 'live_qualcomm':
        extract_frames_residual_yuv(video_path, sampled_path, frame_interval, video_width, video_height, pixfmt, framerate)
    else:
        extract_frames_residual(video_path, sampled_path, frame_interval)
