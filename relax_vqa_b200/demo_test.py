"""Drop-in for src/demo_test.py::evaluate_video_quality.

With ``config['video_path']`` the video itself is sampled (video_frames_extract.sample_clip: raw yuv420p converted on the
GPU, containers decoded by OpenCV's FFmpeg) and goes straight into the engine; without it the entry point reads the PNGs
the reference's sampler has written under ``sampled_root`` (src/demo_test.py:68-69)."""
import glob
import os

import cv2
import numpy as np
import torch
from joblib import load

from . import runtime
from .engine import Clip
from .weights import fix_state_dict   # noqa: F401  (re-exported: src/demo_test.py:25-35)


def _sorted_frames(folder, video_name):
    paths = [p for p in glob.glob(os.path.join(folder, f'{video_name}_*.png'))]
    orig = sorted([p for p in paths if '_next' not in os.path.basename(p) and os.path.basename(p)[len(video_name) + 1:-4].isdigit()],
                  key=lambda x: int(x.split('_')[-1].split('.')[0]))
    nxt = sorted([p for p in paths if os.path.basename(p).endswith('_next.png')], key=lambda x: int(x.split('_')[-2]))
    return orig, nxt


def load_clip_host(sampled_frame_path, sampled_fragment_path, video_name, pin=False):
    """-> (frames [Tf,H,W,3], nexts [Tp,H,W,3]) uint8 BGR host tensors (pinned on request) from the reference's PNG layout."""
    full, _ = _sorted_frames(sampled_frame_path, video_name)
    orig, nxt = _sorted_frames(sampled_fragment_path, video_name)
    if not full:
        raise FileNotFoundError(f"no sampled frames {video_name}_*.png in {sampled_frame_path}")
    n = min(len(orig), len(nxt))

    def rd(ps, like=None):
        imgs = [cv2.imread(p) for p in ps]
        if any(i is None for i in imgs):
            raise ValueError("unreadable PNG among " + ", ".join(os.path.basename(p) for p in ps[:3]))
        if not imgs:
            return torch.empty((0,) + tuple(like.shape[1:]), dtype=torch.uint8)
        t = torch.from_numpy(np.stack(imgs))
        return t.pin_memory() if pin and torch.cuda.is_available() else t

    # the full-frame blocks average over all sampled frames, the fragment blocks over pairs (ref :80-87, :104)
    if [os.path.basename(p) for p in full[:n]] != [os.path.basename(p) for p in orig[:n]]:
        raise ValueError("sampled-frame and fragment folders disagree")
    frames = rd(full)
    return frames, rd(nxt[:n], like=frames)


def load_clip(sampled_frame_path, sampled_fragment_path, video_name, device):
    frames, nexts = load_clip_host(sampled_frame_path, sampled_fragment_path, video_name)
    return Clip(frames.to(device), nexts.to(device))


def evaluate_video_quality(config):
    """ref :51-219.  Same config keys; returns the predicted score (float)."""
    eng = runtime.engine()
    video_type, video_name = config['video_type'], config['video_name']
    save_path = config['save_path']
    if config.get('video_path'):
        from .video_frames_extract import sample_clip
        clip = sample_clip(video_type, config['video_path'], int(config['framerate'] / 2), config.get('video_width'),
                           config.get('video_height'), config.get('pixfmt', 'yuv420p'), eng.device)        # ref :76, :92
    else:
        base = config.get('sampled_root', "../video_sampled_frame/original_sampled_frame/")
        clip = load_clip(os.path.join(base, "test_sampled_frames"), os.path.join(base, "test_sampled_fragment"), video_name, eng.device)
    imputer = load(f'{save_path}/scaler/{video_type}_imputer.pkl')
    scaler = load(f'{save_path}/scaler/{video_type}_scaler.pkl')
    if config['is_finetune'] is True:
        model_path = os.path.join(save_path, f"fine_tune_model/{video_type}_relaxvqa_{config['select_criteria']}_fine_tuned_model.pth")
    else:
        model_path = os.path.join(save_path, f"{config['train_data_name']}_relaxvqa_{config['select_criteria']}_trained_median_model_param_onLSVQ_TEST.pth")
    state_dict = torch.load(model_path, map_location='cpu')
    eng.load_head(state_dict, imputer.statistics_, scaler.scale_, scaler.min_)
    _, score = eng.predict([clip], video_type, is_finetune=config['is_finetune'] is True)
    return float(score[0].item())
