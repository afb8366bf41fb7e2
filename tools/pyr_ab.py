"""A/B of the Farneback path against its environment switches, each variant in its own process (the switches are read
once): default vs B200VQA_PYR_TILE (per-level tile pyramid kernel instead of the fused streaming one), B200VQA_UPSAMPLE_FUSED
(upsampling inside the first iteration of a level instead of the separate flow-upsample pass), ...  Flows are compared bit for bit and the
whole farneback call is timed.
python tools/pyr_ab.py [--height 1080 --width 1920 --pairs 22 --envs B200VQA_PYR_TILE,B200VQA_UPSAMPLE_FUSED]"""
import argparse
import os
import subprocess
import sys

import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--pairs", type=int, default=22)
ap.add_argument("--child", default="")
ap.add_argument("--envs", default="B200VQA_PYR_TILE,B200VQA_UPSAMPLE_FUSED")
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if args.child:
    sys.path.insert(0, ROOT)
    import torch
    from relax_vqa_b200 import ops, synth
    from oracle import fragments as F
    ctx = ops.Context(0)
    fr, nx = synth.make_clip(7, args.height, args.width, args.pairs)
    g0 = torch.from_numpy(np.stack([F.bgr2gray(f) for f in fr])).cuda()
    g1 = torch.from_numpy(np.stack([F.bgr2gray(f) for f in nx])).cuda()
    for _ in range(3):
        flow = ops.farneback(ctx, g0, g1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        flow = ops.farneback(ctx, g0, g1)
    e1.record()
    torch.cuda.synchronize()
    print(f"{args.child}: farneback {e0.elapsed_time(e1) / 5:.3f} ms/call", flush=True)
    np.save(args.child, flow.cpu().numpy())
    ctx.close()
    sys.exit(0)
outs = {}
variants = [("default", {})] + [(e, {e: "1"}) for e in args.envs.split(",") if e]
for name, env in variants:
    path = f"/tmp/pyr_ab_{name}.npy"
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, __file__, "--child", path, "--height", str(args.height), "--width", str(args.width),
                        "--pairs", str(args.pairs)], env=e, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-400:])
    outs[name] = np.load(path)
for name, _ in variants[1:]:
    d = np.abs(outs["default"] - outs[name])
    print(f"{args.height}x{args.width}: max |flow_default - flow[{name}]| = {d.max():.3e} px, mean {d.mean():.3e} px")
