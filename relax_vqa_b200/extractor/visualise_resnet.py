"""Drop-in for src/extractor/visualise_resnet.py::process_video_frame (15-hook layer stack).

The reference runs one full ResNet-50 forward per hooked layer and returns the raw activation maps.  Here one
forward produces all 15 hooks.  Fast path: their spatial means come out of the convolution epilogues, so the
returned mapping holds per-layer pooled vectors (C,); with runtime.configure(return_maps=True) it holds the
reference's (C, H, W) float32 maps (b200vqa_resnet50_maps)."""
from collections import OrderedDict

import numpy as np

from .. import main_fragment_layerstack as _mfl
from .. import runtime

_WIDTHS = [64, 256, 256, 256, 512, 512, 512, 512, 1024, 1024, 1024, 1024, 2048, 2048, 2048]


def process_video_frame(video_name, image_path, all_layers, qp):
    """ref :62-109 -> (OrderedDict layer_name -> activation, frame_npy_path)."""
    _, _, feat = _mfl.get_deep_feature('resnet50', video_name, image_path, qp, 'layer_stack')
    out, o = OrderedDict(), 0
    for name, w in zip(_mfl.RESNET_LAYERS, _WIDTHS):
        if name in all_layers:
            out[name] = feat[name] if runtime.return_maps() else np.asarray(feat[o:o + w])
        o += w
    return out, f'../features/resnet50/{video_name}/frame_{qp}.npy'
