"""Process-wide default Engine used by the reference-named module functions.

The reference creates module-global models at import time on ``cuda`` if available
(src/extractor/visualise_resnet.py:17-21) and always loads pretrained torchvision / DINO weights; here the
engine is created lazily on first use, there is no CPU branch, and ``configure()`` must install the
backbone state dicts first.  Without them ``engine()`` raises: a score computed from random backbones looks
plausible and means nothing.  Seeded synthetic weights (tests, benchmark, smoke; the pretrained files
cannot be downloaded offline) are an explicit opt-in: ``configure(allow_seeded_weights=True)``."""
import numpy as np
import torch

_engine = None
_config = dict(device=0, resnet_sd=None, vit_sd=None, return_maps=False, allow_seeded_weights=False)


def configure(device=0, resnet_sd=None, vit_sd=None, return_maps=False, allow_seeded_weights=False):
    """return_maps=True: get_deep_feature / process_video_frame return the reference's RAW types
    (dict[str, (C,H,W)], (2048,1,1), (196,768)) from the un-fused boundary-fidelity path instead of
    vectors pooled inside the kernels (PooledFrame); process_video_feature accepts both."""
    global _engine
    _config.update(device=device, resnet_sd=resnet_sd, vit_sd=vit_sd, return_maps=bool(return_maps),
                   allow_seeded_weights=bool(allow_seeded_weights))
    if _engine is not None:
        _engine.close()
        _engine = None


def return_maps():
    return _config["return_maps"]


def engine():
    global _engine
    if _engine is None:
        from . import _lib
        from .engine import Engine
        seeded = _config["allow_seeded_weights"]
        if not seeded and (_config["resnet_sd"] is None or _config["vit_sd"] is None):
            raise _lib.B200VQAError(
                "weights not loaded (B200VQA_ENOTLOADED): call relax_vqa_b200.runtime.configure(resnet_sd=..., vit_sd=...) with the "
                "torchvision ResNet-50 and DINO ViT-B/16 state dicts, or configure(allow_seeded_weights=True) for synthetic weights")
        _engine = Engine(_config["device"], _config["resnet_sd"], _config["vit_sd"], seed_if_missing=seeded)
    return _engine


def to_dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(engine().device)


class PooledFrame(np.ndarray):
    """Per-image feature vector whose spatial pooling already happened on the GPU (the reference returns
    raw activation maps here and pools them on the host in process_video_feature)."""

    def __new__(cls, arr, kind):
        obj = np.asarray(arr, dtype=np.float32).view(cls)
        obj.kind = kind
        return obj

    def __array_finalize__(self, obj):
        self.kind = getattr(obj, "kind", None)
