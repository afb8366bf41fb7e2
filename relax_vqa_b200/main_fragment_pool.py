"""Drop-in for src/main_fragment_pool.py (ViT-B/16 pooled features on fragments)."""
import numpy as np

from . import main_fragment_layerstack as _mfl
from .main_fragment_layerstack import (flow_to_rgb, get_patch_diff, extract_important_patches,   # noqa: F401
                                       get_original_frame_patches, process_patches, merge_fragments,
                                       concatenate_features)


def pool_vit_tokens(frame):
    """ref :124-132: hstack([mean, max, std]) over the 196 tokens (ddof = 0) -> (2304,).  A PooledFrame (fast path:
    pooled by k11_final_norm_pool) passes through; a raw (196, 768) array (return_maps=True) is pooled here."""
    a = np.asarray(frame, dtype=np.float32)
    if a.ndim == 2:
        return np.hstack([np.mean(a, axis=0), np.max(a, axis=0), np.std(a, axis=0)]).astype(np.float32)
    return a


def get_deep_feature(network_name, video_name, image_path, qp, layer_name):
    """ref :83-111."""
    return _mfl.get_deep_feature(network_name, video_name, image_path, qp, layer_name)


def process_video_feature(video_feature, network_name):
    """ref :114-143 -> (T, 2304)."""
    return np.array([pool_vit_tokens(f) for f in video_feature])
