"""One profiled step of the hot path (for ncu): warm up, then run `--steps` steps between
cudaProfilerStart/Stop.  Use with: ncu --profile-from-start off ... python tools/profile_step.py"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from relax_vqa_b200 import weights
from relax_vqa_b200.engine import Engine, synthetic_clips_on_device

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=2)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--pairs", type=int, default=22)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
eng = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
clips = synthetic_clips_on_device(a.clips, a.height, a.width, a.pairs, eng.device, seed=5)
for _ in range(a.warmup):
    eng.predict(clips, "live_vqc")
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    eng.predict(clips, "live_vqc")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", eng.launches)
