"""Oracle (CPU, torch fp32 functional ops) for A10-A15: backbone activations and pooling.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Parity for these rows is "unpinned" by reference artefacts (no stored features); the
restatement is pinned against the reference itself, imported through shims with the same
seeded weights (tests/golden/gen_golden.py, tests/test_oracle_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

RESNET_MEAN = (0.485, 0.456, 0.406)
RESNET_STD = (0.229, 0.224, 0.225)
STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))
HOOKED_BLOCKS = (3, 4, 4, 3)     # src/main_fragment_layerstack.py:91-95 (layer3[4], [5] not hooked)


def resnet_preprocess(rgb_u8):
    """ToTensor + Normalize - src/extractor/visualise_resnet.py:40-50. (B,224,224,3) u8 RGB."""
    x = torch.from_numpy(np.ascontiguousarray(rgb_u8)).permute(0, 3, 1, 2).float().div(255.0)
    mean = torch.tensor(RESNET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(RESNET_STD).view(1, 3, 1, 1)
    return (x - mean) / std


def vit_preprocess(rgb_u8):
    """ToTensor only - src/extractor/visualise_vit_layer.py:339-342."""
    return torch.from_numpy(np.ascontiguousarray(rgb_u8)).permute(0, 3, 1, 2).float().div(255.0)


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


@torch.no_grad()
def resnet50_activations(sd, x):
    """torchvision ResNet-50 v1.5 forward; returns the 15 hooked maps in hook order
    (src/extractor/visualise_resnet.py:83-106) plus the avgpool output (B,2048)."""
    acts = []
    y = F.conv2d(x, sd["conv1.weight"], stride=2, padding=3)
    acts.append(y)                                    # hooked raw, before bn1/relu
    y = F.max_pool2d(F.relu(_bn(y, sd, "bn1")), 3, 2, 1)
    for li, (planes, blocks, stride) in enumerate(STAGES, start=1):
        for b in range(blocks):
            p = f"layer{li}.{b}"
            s = stride if b == 0 else 1
            idt = y
            o = F.relu(_bn(F.conv2d(y, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
            o = F.relu(_bn(F.conv2d(o, sd[p + ".conv2.weight"], stride=s, padding=1), sd, p + ".bn2"))
            o = _bn(F.conv2d(o, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            if b == 0:
                idt = _bn(F.conv2d(y, sd[p + ".downsample.0.weight"], stride=s), sd, p + ".downsample.1")
            y = F.relu(o + idt)
            if b < HOOKED_BLOCKS[li - 1]:
                acts.append(y)
    pooled = y.mean(dim=(2, 3))
    return acts, pooled


def resnet50_layerstack(sd, x):
    """A12: per-hook spatial mean, concatenated -> (B, 13120) fp32
    (process_video_feature 'layer_stack', src/main_fragment_layerstack.py:131-140)."""
    acts, _ = resnet50_activations(sd, x)
    return np.concatenate([a.numpy().mean(axis=(2, 3)) for a in acts], axis=1)


def resnet50_pool(sd, x):
    """A13: avgpool vector + [mean, max, std] scalars -> (B, 2051)
    (src/main_fragment_layerstack.py:141-149)."""
    _, pooled = resnet50_activations(sd, x)
    v = pooled.numpy()
    return np.concatenate([v, v.mean(1, keepdims=True), v.max(1, keepdims=True), v.std(1, keepdims=True)], axis=1)


@torch.no_grad()
def vit_tokens(sd, x, depth=12, heads=12):
    """A14: DINO ViT-B/16 -> final-LayerNorm patch tokens (B,196,768)
    (src/extractor/visualise_vit_layer.py:152-260, :492-500)."""
    B = x.shape[0]
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=16)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], dim=1) + sd["pos_embed"]
    D = t.shape[-1]
    hd = D // heads
    for i in range(depth):
        p = f"blocks.{i}"
        h = F.layer_norm(t, (D,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps=1e-6)
        qkv = F.linear(h, sd[p + ".attn.qkv.weight"], sd[p + ".attn.qkv.bias"])
        qkv = qkv.reshape(B, -1, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * (hd ** -0.5)).softmax(dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(B, -1, D)
        t = t + F.linear(o, sd[p + ".attn.proj.weight"], sd[p + ".attn.proj.bias"])
        h = F.layer_norm(t, (D,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps=1e-6)
        h = F.gelu(F.linear(h, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"]))
        t = t + F.linear(h, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
    t = F.layer_norm(t, (D,), sd["norm.weight"], sd["norm.bias"], eps=1e-6)
    return t[:, 1:]


def vit_pool(sd, x):
    """A15: [mean, max, std(ddof=0)] over the 196 tokens -> (B, 2304)
    (src/main_fragment_pool.py:124-132)."""
    tok = vit_tokens(sd, x).numpy()
    return np.concatenate([tok.mean(1), tok.max(1), tok.std(1)], axis=1)
