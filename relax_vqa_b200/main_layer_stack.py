"""Drop-in for src/main_layer_stack.py (full-frame ResNet-50 layer-stack / ViT pooled features)."""
import numpy as np

from . import main_fragment_layerstack as _mfl
from .main_fragment_pool import pool_vit_tokens


def get_deep_feature(network_name, video_name, image_path, qp):
    """ref :81-112 (4-argument variant)."""
    layer = 'layer_stack' if network_name == 'resnet50' else 'pool'
    return _mfl.get_deep_feature(network_name, video_name, image_path, qp, layer)


def process_video_feature(video_feature, network_name):
    """ref :115-151 -> (T, 13120) for resnet50 (per-layer spatial mean, :138), (T, 2304) for vit (:126-131).
    Accepts pooled vectors (fast path) and the reference's raw maps / tokens (return_maps=True)."""
    if network_name == 'resnet50':
        return _mfl.process_video_feature(video_feature, network_name, 'layer_stack')
    return np.array([pool_vit_tokens(f) for f in video_feature])
