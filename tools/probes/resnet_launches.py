"""One ResNet-50 pass of 264 images (for ncu --metrics gpu__time_duration.sum -k regex:gemm_tcgen05)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from relax_vqa_b200 import ops, weights
from relax_vqa_b200.engine import Engine
eng = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
img = torch.randint(0, 256, (264, 224, 224, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    ops.resnet50_features(eng.ctx, img, is_bgr=True, want_stack=True, want_pool=True)
torch.cuda.synchronize()
