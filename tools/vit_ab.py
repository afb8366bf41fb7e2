"""Times the ViT-B/16 (and ResNet-50) feature passes alone (CUDA events): images per call = --images.
Use B200VQA_GEMM_NOEPI=8 to switch the fp16 linears back to the staged epilogue (A/B)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relax_vqa_b200 import ops, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=127)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
ctx = ops.Context(0)
ops.load_vitb16(ctx, weights.seeded_vitb16_state_dict())
ops.load_resnet50(ctx, weights.seeded_resnet50_state_dict())
img = torch.randint(0, 256, (args.images, 224, 224, 3), dtype=torch.uint8, device="cuda")
for name, fn in (("vit", lambda: ops.vitb16_features(ctx, img)), ("resnet", lambda: ops.resnet50_features(ctx, img))):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    ctx.set_profiling(True); ctx.profile_read()
    for _ in range(args.reps):
        fn()
    gms, n, fl = ctx.profile_read()
    ctx.set_profiling(False)
    o = out[0] if isinstance(out, (tuple, list)) else out
    print(f"{name}: {ms:.3f} ms per {args.images} images; tcgen05 launches {gms / args.reps:.3f} ms ({n // args.reps} launches, "
          f"{fl / (gms / 1e3) / 1e12:.0f} TFLOP/s); checksum {float(o.double().sum()):.6f}", flush=True)
ctx.close()
