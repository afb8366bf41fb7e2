"""Drop-in for src/main_fragment_pool.py (ViT-B/16 pooled features on fragments)."""
import numpy as np

from . import main_fragment_layerstack as _mfl
from .main_fragment_layerstack import (flow_to_rgb, get_patch_diff, extract_important_patches,   # noqa: F401
                                       get_original_frame_patches, process_patches, merge_fragments,
                                       concatenate_features)


def get_deep_feature(network_name, video_name, image_path, qp, layer_name):
    """ref :83-111."""
    return _mfl.get_deep_feature(network_name, video_name, image_path, qp, layer_name)


def process_video_feature(video_feature, network_name):
    """ref :114-143 -> (T, 2304)."""
    return np.array([np.asarray(f, dtype=np.float32) for f in video_feature])
