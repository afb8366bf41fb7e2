"""A/B of the engine's scheduling knobs on the 1080p workload: software pipeline across calls on / off, SMs left to the
image stages (gemm_sms), number of lanes.  Prints videos/s per setting (CUDA events, device-resident inputs)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relax_vqa_b200 import weights  # noqa: E402
from relax_vqa_b200.engine import Engine, synthetic_clips_on_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=4)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--hw", default="1080x1920")
ap.add_argument("--pairs", type=int, default=22)
ap.add_argument("--sms", default="0,140,132,124,116,108")
a = ap.parse_args()
H, W = (int(v) for v in a.hw.split("x"))
eng = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
clips = synthetic_clips_on_device(a.clips, H, W, a.pairs, eng.device, seed=1000)
stream = torch.cuda.current_stream()


def run(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        eng.predict(clips, "live_vqc")
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


ref = None
for pipeline in (False, True):
    for sms in [int(v) for v in a.sms.split(",")]:
        if not pipeline and sms:
            continue
        eng.pipeline = pipeline
        eng.set_gemm_sms(sms)
        run(3)
        ms = min(run(a.steps) for _ in range(2))
        f, s = eng.predict(clips, "live_vqc")
        torch.cuda.synchronize()
        if ref is None:
            ref = (f.clone(), s.clone())
        same = torch.equal(ref[0], f) and torch.equal(ref[1], s)
        print(f"pipeline={int(pipeline)} gemm_sms={sms:3d}: {a.clips * a.steps / ms * 1e3:7.2f} videos/s  ({ms / a.steps:6.2f} ms/step)  identical={same}", flush=True)
