"""A/B of Engine.GROUP_PIXELS (fusing the image stages of same-resolution clips into one call) and LANES on a workload."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from relax_vqa_b200 import weights  # noqa: E402
from relax_vqa_b200.engine import Engine, synthetic_clips_on_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="1080p-10s")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--groups", default="0,60e6,120e6,240e6")
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
specs = bench.global_specs(a.workload, wl["clips"])
eng = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
clips = []
for gi, (h, w, p) in enumerate(specs):
    clips += synthetic_clips_on_device(1, h, w, p, eng.device, seed=1000 + gi)
stream = torch.cuda.current_stream()
ref = None
for g in [int(float(v)) for v in a.groups.split(",")]:
    eng.GROUP_PIXELS = g
    for _ in range(3):
        eng.predict(clips, wl["video_type"])
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(a.steps):
            f, s = eng.predict(clips, wl["video_type"])
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    if ref is None:
        ref = (f.clone(), s.clone())
    print(f"{a.workload} GROUP_PIXELS={g:>10d}: {len(clips) / best * 1e3:7.2f} videos/s ({best:7.2f} ms/step) units={len(eng._units(clips))} "
          f"identical={torch.equal(ref[0], f) and torch.equal(ref[1], s)}", flush=True)
