"""A/B of the ViT attention kernels: tcgen05 / TMEM (attn_impl 0) vs warp-level mma.sync (attn_impl 1) vs the SIMT check
kernel (gemm_impl 1).  Prints the ViT pass time per setting (the 12 attention launches are the only difference between 0
and 1) and the agreement of the pooled features."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relax_vqa_b200 import ops, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=264)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
ctx = ops.Context(0)
ops.load_vitb16(ctx, weights.seeded_vitb16_state_dict())
img = torch.randint(0, 256, (a.images, 224, 224, 3), dtype=torch.uint8, device="cuda")
res = {}
for impl in (1, 2, 0):
    ctx.set_attn_impl(impl)
    for _ in range(3):
        out = ops.vitb16_features(ctx, img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = ops.vitb16_features(ctx, img)
    e1.record()
    torch.cuda.synchronize()
    res[impl] = (e0.elapsed_time(e1) / a.reps, out.clone())
    print(f"attn_impl={impl}: ViT pass {res[impl][0]:.3f} ms per {a.images} images", flush=True)
small = img[:6].contiguous()
ctx.set_gemm_impl(1)
chk = ops.vitb16_features(ctx, small)
ctx.set_gemm_impl(0)
for impl in (1, 2, 0):
    ctx.set_attn_impl(impl)
    got = ops.vitb16_features(ctx, small)
    err = (got - chk).abs().max().item() / chk.pow(2).mean().sqrt().item()
    print(f"attn_impl={impl} vs SIMT check: max|d|/rms = {err:.2e}")
d = (res[0][1] - res[1][1]).abs().max().item() / res[1][1].pow(2).mean().sqrt().item()
print(f"tcgen05 vs mma.sync attention: max|d|/rms = {d:.2e}; saving per ViT pass {res[1][0] - res[0][0]:.3f} ms "
      f"= {(res[1][0] - res[0][0]) / 12 * 1e3:.1f} us per attention launch (2-CTA/SM version: {(res[1][0] - res[2][0]) / 12 * 1e3:.1f} us)")
ctx.close()
