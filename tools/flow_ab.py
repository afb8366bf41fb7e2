"""A/B timing of the Farneback iteration kernels (CUDA events, 1080p x 22 pairs by default).
python tools/flow_ab.py [--height 1080 --width 1920 --pairs 22 --static]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relax_vqa_b200 import ops, synth  # noqa: E402
from oracle import fragments as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--pairs", type=int, default=22)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--impls", default="0,5,4,1,2")
ap.add_argument("--static", action="store_true", help="static scene + noise (near-zero flow: worst case for tap reuse)")
args = ap.parse_args()

ctx = ops.Context(0)
fr, nx = synth.make_clip(7, args.height, args.width, args.pairs)
if args.static:
    nx = np.clip(fr.astype(np.int16) + np.random.default_rng(1).integers(-2, 3, fr.shape), 0, 255).astype(np.uint8)
g0 = torch.from_numpy(np.stack([F.bgr2gray(f) for f in fr])).cuda()
g1 = torch.from_numpy(np.stack([F.bgr2gray(f) for f in nx])).cuda()
ref = None
for impl in [int(v) for v in args.impls.split(",")]:
    ctx.set_flow_impl(impl)
    for _ in range(2):
        flow = ops.farneback(ctx, g0, g1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        flow = ops.farneback(ctx, g0, g1)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / args.reps
    ctx.set_profiling(True)
    ctx.profile_read_flow()
    for _ in range(args.reps):
        ops.farneback(ctx, g0, g1)
    ms, n, by = ctx.profile_read_flow()
    ctx.set_profiling(False)
    if ref is None:
        ref = flow.clone()
    d = (flow - ref).abs()
    print(f"impl {impl}: farneback {total:.3f} ms/call; iteration kernel {ms / args.reps:.3f} ms/call over {n // args.reps} launches "
          f"= {by / (ms / 1e3) / 1e9:.0f} GB/s algorithmic; max |flow - first impl| = {float(d.max()):.2e}", flush=True)
ctx.close()
