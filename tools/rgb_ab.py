"""A/B of the flow-colouring kernels: warp-per-band (default) vs the 64 x 16 block form (B200VQA_RGB_BLOCK=1).
python tools/rgb_ab.py [--height 1080 --width 1920 --pairs 22]"""
import argparse, os, subprocess, sys
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--pairs", type=int, default=22)
ap.add_argument("--child", default="")
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if args.child:
    sys.path.insert(0, ROOT)
    import torch
    from relax_vqa_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    flow = torch.randn(args.pairs, args.height, args.width, 2, device="cuda", generator=g) * 3.0
    flow[0, :40] = 0.0                                   # zero vectors: hue undefined
    for want_rgb in (True, False):
        for _ in range(3):
            rgb, sums, mm = ops.flow_to_rgb(flow, want_rgb=want_rgb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            rgb, sums, mm = ops.flow_to_rgb(flow, want_rgb=want_rgb)
        e1.record(); torch.cuda.synchronize()
        print(f"{os.path.basename(args.child)} want_rgb={want_rgb}: flow_to_rgb {e0.elapsed_time(e1) / 10 * 1e3:.1f} us/call", flush=True)
        if want_rgb:
            np.save(args.child + ".rgb.npy", rgb.cpu().numpy())
    np.save(args.child + ".sums.npy", sums.cpu().numpy())
    sys.exit(0)
res = []
for name, env in (("band", {}), ("block", {"B200VQA_RGB_BLOCK": "1"})):
    path = f"/tmp/rgb_ab_{name}"
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, __file__, "--child", path, "--height", str(args.height), "--width", str(args.width),
                        "--pairs", str(args.pairs)], env=e, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-400:])
    res.append((np.load(path + ".rgb.npy"), np.load(path + ".sums.npy")))
print(f"{args.height}x{args.width}: rgb identical {bool(np.array_equal(res[0][0], res[1][0]))}, sums identical {bool(np.array_equal(res[0][1], res[1][1]))}")
