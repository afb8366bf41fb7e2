"""Timing of the two-filter resize call (both horizontal passes from one staged row) against two single-filter calls.
python tools/resize_ab.py [--height 1080 --width 1920 --frames 22]"""
import argparse, os, subprocess, sys
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--frames", type=int, default=22)
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if True:
    sys.path.insert(0, ROOT)
    import torch
    from relax_vqa_b200 import ops
    ctx = ops.Context(0)
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.randint(0, 256, (args.frames, args.height, args.width, 3), device="cuda", generator=g, dtype=torch.uint8)
    def timed(fn, name):
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us/call", flush=True)
        return out
    pair = timed(lambda: ops.resize_pil_pair(ctx, img), "resize_pil_pair")
    two = timed(lambda: (ops.resize_pil(ctx, img, 0), ops.resize_pil(ctx, img, 1)), "two resize_pil calls")
    print(f"{args.height}x{args.width}: identical {bool(torch.equal(pair[0], two[0]) and torch.equal(pair[1], two[1]))}")
    ctx.close()
    sys.exit(0)
