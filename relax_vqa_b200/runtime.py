"""Process-wide default Engine used by the reference-named module functions.

The reference creates module-global models at import time on ``cuda`` if available
(src/extractor/visualise_resnet.py:17-21); here the engine is created lazily on first use and there is
no CPU branch.  ``configure()`` installs real checkpoints; without it seeded synthetic weights are used
(the pretrained files cannot be downloaded offline)."""
import numpy as np
import torch

_engine = None
_config = dict(device=0, resnet_sd=None, vit_sd=None)


def configure(device=0, resnet_sd=None, vit_sd=None):
    global _engine
    _config.update(device=device, resnet_sd=resnet_sd, vit_sd=vit_sd)
    if _engine is not None:
        _engine.close()
        _engine = None


def engine():
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine(_config["device"], _config["resnet_sd"], _config["vit_sd"])
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
