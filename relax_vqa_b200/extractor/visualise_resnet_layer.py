"""Drop-in for src/extractor/visualise_resnet_layer.py::process_video_frame (avgpool hook)."""
import numpy as np

from .. import main_fragment_layerstack as _mfl


def process_video_frame(video_name, image_path, layer_name, qp):
    """ref :62-102 -> (ndarray (2048, 1, 1) float32, frame_npy_path)."""
    if layer_name != 'resnet50.avgpool':
        raise ValueError("only the 'resnet50.avgpool' hook is on the hot path (src/main_fragment_layerstack.py:98)")
    _, _, vec = _mfl.get_deep_feature('resnet50', video_name, image_path, qp, 'pool')
    return np.asarray(vec).reshape(2048, 1, 1), f'../features/resnet50/{video_name}/frame_{qp}.npy'
