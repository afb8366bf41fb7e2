"""Drop-in for src/extractor/visualise_vit_layer.py: VitGenerator + process_video_frame.

The reference rebuilds and reloads ViT-B/16 for every image (src/main_fragment_layerstack.py:118); here
the weights live in the engine and VitGenerator is a handle to them."""
import numpy as np

from .. import main_fragment_layerstack as _mfl
from .. import runtime


class VitGenerator(object):
    def __init__(self, name_model, patch_size, device, evaluate=True, random=False, verbose=False, state_dict=None):
        if name_model != 'vit_base' or patch_size != 16:
            raise ValueError("the B200 hot path implements DINO ViT-B/16 only (ref :287-289)")
        self.name_model, self.patch_size, self.device = name_model, patch_size, device
        self.evaluate, self.random, self.verbose = evaluate, random, verbose     # kept for signature parity; the engine is always in eval mode
        if random:
            raise ValueError("random=True (an un-initialised ViT) is not supported: weights come from runtime.configure / state_dict")
        if state_dict is not None:
            from .. import ops
            ops.load_vitb16(runtime.engine().ctx, state_dict)


def process_video_frame(image_path, video_name, qp, model, patch_size, device):
    """ref :447-490.  Fast path: returns (pooled (2304,) float32 = [mean|max|std] over the 196 final-LayerNorm patch
    tokens, npy path); with runtime.configure(return_maps=True): the reference's (196, 768) tokens."""
    _, _, vec = _mfl.get_deep_feature('vit', video_name, image_path, qp, 'pool')
    return vec, f'../features/vit/{video_name}/frame_attention_{qp}.npy'
