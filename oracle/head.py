"""Oracle (CPU) for A17/A18: imputer + scaler + Mlp regression head.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.
"""
import numpy as np
import torch
import torch.nn.functional as F


def impute_scale(x, imputer_mean, scale, minv):
    """SimpleImputer(mean).transform then MinMaxScaler.transform (src/demo_test.py:177-180),
    applied from the fitted attributes (statistics_, scale_, min_) in float64."""
    x = np.asarray(x, dtype=np.float64).copy()
    nan = np.isnan(x)
    x[nan] = np.broadcast_to(imputer_mean, x.shape)[nan]
    return x * scale + minv


@torch.no_grad()
def mlp_forward(sd, x):
    """Mlp.forward in eval mode (src/model_regression.py:49-58): fc1 -> BatchNorm1d (running
    stats, eps 1e-5) -> GELU(erf) -> fc2 -> GELU -> fc3.  Dropout is the identity in eval."""
    x = torch.as_tensor(np.asarray(x), dtype=torch.float32)
    h = F.linear(x, sd["fc1.weight"], sd["fc1.bias"])
    h = F.batch_norm(h, sd["bn1.running_mean"], sd["bn1.running_var"], sd["bn1.weight"], sd["bn1.bias"],
                     training=False, eps=1e-5)
    h = F.gelu(h)
    h = F.gelu(F.linear(h, sd["fc2.weight"], sd["fc2.bias"]))
    return F.linear(h, sd["fc3.weight"], sd["fc3.bias"]).squeeze(-1).numpy()


def rescale_score(pred, video_type, is_finetune=False):
    """src/demo_test.py:206-219: LSVQ-trained head on KoNViD-1k / YouTube-UGC -> 1..5 scale."""
    if not is_finetune and video_type in ("youtube_ugc", "konvid_1k"):
        return (pred / 100.0) * 4.0 + 1.0
    return pred
