"""Small ViT / ResNet pass for ncu captures: python tools/profile_vit.py --images 264 [--net vit|resnet] [--attn-impl 0|1]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relax_vqa_b200 import ops, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=264)
ap.add_argument("--net", default="vit")
ap.add_argument("--attn-impl", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
ctx = ops.Context(0)
ctx.set_attn_impl(a.attn_impl)
img = torch.randint(0, 256, (a.images, 224, 224, 3), dtype=torch.uint8, device="cuda")
if a.net == "vit":
    ops.load_vitb16(ctx, weights.seeded_vitb16_state_dict())
    for _ in range(a.reps):
        ops.vitb16_features(ctx, img)
else:
    ops.load_resnet50(ctx, weights.seeded_resnet50_state_dict())
    for _ in range(a.reps):
        ops.resnet50_features(ctx, img)
torch.cuda.synchronize()
ctx.close()
