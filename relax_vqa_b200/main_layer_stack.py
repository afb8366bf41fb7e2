"""Drop-in for src/main_layer_stack.py (full-frame ResNet-50 layer-stack / ViT pooled features)."""
import numpy as np

from . import main_fragment_layerstack as _mfl


def get_deep_feature(network_name, video_name, image_path, qp):
    """ref :81-112 (4-argument variant)."""
    layer = 'layer_stack' if network_name == 'resnet50' else 'pool'
    return _mfl.get_deep_feature(network_name, video_name, image_path, qp, layer)


def process_video_feature(video_feature, network_name):
    """ref :115-151 -> (T, 13120) for resnet50, (T, 2304) for vit."""
    return np.array([np.asarray(f, dtype=np.float32) for f in video_feature])
