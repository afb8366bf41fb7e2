"""Backbone / head parameter tables and seeded synthetic parameters.

Pretrained checkpoints (torchvision IMAGENET1K_V1 ResNet-50, DINO ViT-B/16, the
*_trained_median_model_param.pth heads) are not available offline, so tests and the
benchmark use seeded synthetic parameters.  They are drawn from numpy's PCG64 stream
(bit-reproducible across machines, unlike torch's vectorised CPU normal_()), keyed and
shaped exactly like the state dicts the reference loads:

  * ResNet-50: torchvision ``models.resnet50`` keys (src/extractor/visualise_resnet.py:21)
  * ViT-B/16 : DINO ``VisionTransformer`` keys (src/extractor/visualise_vit_layer.py:152-260)
  * head     : ``Mlp`` keys (src/model_regression.py:37-47), optionally wrapped the way
               torch.optim.swa_utils.AveragedModel saves them (``module.`` prefix +
               ``n_averaged``), which ``fix_state_dict`` (src/demo_test.py:25-35) strips.

BatchNorm statistics are deliberately non-trivial (and the last BN of every bottleneck is
damped) so that activations stay O(1) through the 16 residual blocks, as they do with
trained weights, and so that a wrong scale/shift index cannot hide behind an identity BN.
"""
from collections import OrderedDict

import numpy as np
import torch

RESNET_LAYERS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))   # (planes, blocks, stride)
RESNET_HOOK_BLOCKS = (3, 4, 4, 3)          # blocks hooked per stage (src/main_fragment_layerstack.py:91-95)
RESNET_STACK_DIM = 64 + 3 * 256 + 4 * 512 + 4 * 1024 + 3 * 2048   # 13120
VIT_DIM, VIT_DEPTH, VIT_HEADS, VIT_TOKENS = 768, 12, 12, 197
FEATURE_DIM = 35203


def resnet50_spec():
    """Ordered (key, shape) list of torchvision's resnet50 state dict."""
    spec = []

    def bn(prefix, c):
        spec.extend([(f"{prefix}.weight", (c,)), (f"{prefix}.bias", (c,)),
                     (f"{prefix}.running_mean", (c,)), (f"{prefix}.running_var", (c,)),
                     (f"{prefix}.num_batches_tracked", ())])

    spec.append(("conv1.weight", (64, 3, 7, 7)))
    bn("bn1", 64)
    inplanes = 64
    for li, (planes, blocks, _stride) in enumerate(RESNET_LAYERS, start=1):
        for b in range(blocks):
            p = f"layer{li}.{b}"
            spec.append((f"{p}.conv1.weight", (planes, inplanes, 1, 1)))
            bn(f"{p}.bn1", planes)
            spec.append((f"{p}.conv2.weight", (planes, planes, 3, 3)))
            bn(f"{p}.bn2", planes)
            spec.append((f"{p}.conv3.weight", (planes * 4, planes, 1, 1)))
            bn(f"{p}.bn3", planes * 4)
            if b == 0:
                spec.append((f"{p}.downsample.0.weight", (planes * 4, inplanes, 1, 1)))
                bn(f"{p}.downsample.1", planes * 4)
            inplanes = planes * 4
    spec.append(("fc.weight", (1000, 2048)))
    spec.append(("fc.bias", (1000,)))
    return spec


def vitb16_spec():
    """Ordered (key, shape) list of the DINO ViT-B/16 state dict."""
    d = VIT_DIM
    spec = [("cls_token", (1, 1, d)), ("pos_embed", (1, VIT_TOKENS, d)),
            ("patch_embed.proj.weight", (d, 3, 16, 16)), ("patch_embed.proj.bias", (d,))]
    for i in range(VIT_DEPTH):
        p = f"blocks.{i}"
        spec += [(f"{p}.norm1.weight", (d,)), (f"{p}.norm1.bias", (d,)),
                 (f"{p}.attn.qkv.weight", (3 * d, d)), (f"{p}.attn.qkv.bias", (3 * d,)),
                 (f"{p}.attn.proj.weight", (d, d)), (f"{p}.attn.proj.bias", (d,)),
                 (f"{p}.norm2.weight", (d,)), (f"{p}.norm2.bias", (d,)),
                 (f"{p}.mlp.fc1.weight", (4 * d, d)), (f"{p}.mlp.fc1.bias", (4 * d,)),
                 (f"{p}.mlp.fc2.weight", (d, 4 * d)), (f"{p}.mlp.fc2.bias", (d,))]
    spec += [("norm.weight", (d,)), ("norm.bias", (d,))]
    return spec


def head_spec(input_features=FEATURE_DIM, hidden=256):
    return [("fc1.weight", (hidden, input_features)), ("fc1.bias", (hidden,)),
            ("bn1.weight", (hidden,)), ("bn1.bias", (hidden,)),
            ("bn1.running_mean", (hidden,)), ("bn1.running_var", (hidden,)),
            ("bn1.num_batches_tracked", ()),
            ("fc2.weight", (hidden // 2, hidden)), ("fc2.bias", (hidden // 2,)),
            ("fc3.weight", (1, hidden // 2)), ("fc3.bias", (1,))]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def seeded_resnet50_state_dict(seed=1234):
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for key, shape in resnet50_spec():
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(1, dtype=torch.long)
        elif key.endswith("conv1.weight") or key.endswith("conv2.weight") or key.endswith("conv3.weight") \
                or key.endswith("downsample.0.weight"):
            fan_out = shape[0] * shape[2] * shape[3]
            sd[key] = _t(rng.standard_normal(shape) * np.sqrt(2.0 / fan_out))
        elif key.endswith("running_mean"):
            sd[key] = _t(rng.standard_normal(shape) * 0.1)
        elif key.endswith("running_var"):
            sd[key] = _t(rng.uniform(0.6, 1.4, shape))
        elif key.endswith(".weight") and len(shape) == 1:            # BN gamma
            damp = 0.35 if (".bn3." in key) else 1.0
            sd[key] = _t(rng.uniform(0.7, 1.2, shape) * damp)
        elif key.endswith(".bias") and len(shape) == 1 and not key.startswith("fc."):   # BN beta
            sd[key] = _t(rng.standard_normal(shape) * 0.1)
        elif key == "fc.weight":
            sd[key] = _t(rng.uniform(-1, 1, shape) / np.sqrt(shape[1]))
        elif key == "fc.bias":
            sd[key] = _t(rng.uniform(-1, 1, shape) / np.sqrt(2048))
        else:
            raise KeyError(key)
    return sd


def seeded_vitb16_state_dict(seed=4321):
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for key, shape in vitb16_spec():
        if key in ("cls_token", "pos_embed"):
            sd[key] = _t(rng.standard_normal(shape) * 0.2)
        elif key == "patch_embed.proj.weight":
            sd[key] = _t(rng.standard_normal(shape) * (2.0 / np.sqrt(768)))
        elif "norm" in key and key.endswith(".weight"):
            sd[key] = _t(rng.uniform(0.8, 1.2, shape))
        elif "norm" in key and key.endswith(".bias"):
            sd[key] = _t(rng.standard_normal(shape) * 0.05)
        elif key.endswith(".weight"):
            sd[key] = _t(rng.standard_normal(shape) * (0.8 / np.sqrt(shape[1])))
        elif key.endswith(".bias"):
            sd[key] = _t(rng.standard_normal(shape) * 0.02)
        else:
            raise KeyError(key)
    return sd


def seeded_head_state_dict(seed=99, input_features=FEATURE_DIM, swa_format=False):
    """Seeded ``Mlp`` parameters; ``swa_format`` mimics AveragedModel.state_dict()
    (src/model_regression.py:388,715)."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for key, shape in head_spec(input_features):
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(7, dtype=torch.long)
        elif key == "bn1.running_var":
            sd[key] = _t(rng.uniform(0.5, 2.0, shape))
        elif key == "bn1.running_mean":
            sd[key] = _t(rng.standard_normal(shape) * 0.3)
        elif key == "bn1.weight":
            sd[key] = _t(rng.uniform(0.8, 1.2, shape))
        elif key.endswith(".weight"):
            sd[key] = _t(rng.uniform(-1, 1, shape) * (3.0 / np.sqrt(shape[1])))
        else:
            sd[key] = _t(rng.uniform(-1, 1, shape) * 0.5)
    sd["fc3.bias"] = sd["fc3.bias"] + 50.0            # keep scores in the 0-100 MOS range
    if swa_format:
        out = OrderedDict()
        out["n_averaged"] = torch.tensor(5, dtype=torch.long)
        for k, v in sd.items():
            out["module." + k] = v
        return out
    return sd


def fix_state_dict(state_dict):
    """Strip AveragedModel's ``module.`` prefix / ``n_averaged`` - same contract as the
    reference helper (src/demo_test.py:25-35, src/fine_tune.py:99-109)."""
    out = OrderedDict()
    for k, v in state_dict.items():
        if k == "n_averaged":
            continue
        out[k[7:] if k.startswith("module.") else k] = v
    return out
