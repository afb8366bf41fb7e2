"""Tiny end-to-end pass for compute-sanitizer (memcheck): every kernel once, small shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relax_vqa_b200 import synth, weights
from relax_vqa_b200.engine import Clip, Engine

eng = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
for hw in ((144, 256), (100, 150)):
    fr, nx = synth.make_clip(1, hw[0], hw[1], 2)
    feats, score = eng.predict([Clip(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda())], "konvid_1k")
    torch.cuda.synchronize()
    print(hw, float(score[0]), bool(torch.isfinite(feats).all()))
eng.close()
