"""Drop-in for src/extractor/visualise_resnet.py::process_video_frame (15-hook layer stack).

The reference runs one full ResNet-50 forward per hooked layer and returns the raw activation maps; here
one forward produces all 15 spatial means inside the convolution epilogues, so the returned mapping holds
the per-layer pooled vectors (C,) instead of (C, H, W) maps."""
from collections import OrderedDict

import numpy as np

from .. import main_fragment_layerstack as _mfl

_WIDTHS = [64, 256, 256, 256, 512, 512, 512, 512, 1024, 1024, 1024, 1024, 2048, 2048, 2048]


def process_video_frame(video_name, image_path, all_layers, qp):
    """ref :62-109 -> (OrderedDict layer_name -> pooled (C,) float32, frame_npy_path)."""
    _, _, vec = _mfl.get_deep_feature('resnet50', video_name, image_path, qp, 'layer_stack')
    out, o = OrderedDict(), 0
    for name, w in zip(_mfl.RESNET_LAYERS, _WIDTHS):
        if name in all_layers:
            out[name] = np.asarray(vec[o:o + w])
        o += w
    return out, f'../features/resnet50/{video_name}/frame_{qp}.npy'
