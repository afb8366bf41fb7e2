"""Tiny end-to-end pass for compute-sanitizer (memcheck / racecheck): every kernel once, small shapes - the inference path
(two resolutions), the raw-map / token boundary path, the yuv sampler and two trainer steps."""
import os, sys
os.environ.setdefault("B200VQA_PYR_OS3", "8")      # small frames still go through the fused pyramid kernel
os.environ.setdefault("B200VQA_RGB_BAND", "1")     # ... and the warp-per-band colouring kernel
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from relax_vqa_b200 import ops, synth, weights
from relax_vqa_b200.engine import Clip, Engine
from relax_vqa_b200.model_regression import HeadTrainer

eng = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
for hw in ((144, 256), (100, 150)):
    fr, nx = synth.make_clip(1, hw[0], hw[1], 2)
    feats, score = eng.predict([Clip(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda())], "konvid_1k")
    torch.cuda.synchronize()
    print(hw, float(score[0]), bool(torch.isfinite(feats).all()))
g0 = torch.randint(0, 256, (2, 272, 480), dtype=torch.uint8, device="cuda")
flow, fsums, mm = ops.farneback_flow_sums(eng.ctx, g0, torch.roll(g0, 2, 2))
A = (torch.randn(300, 128, device="cuda") * 0.5).half(); Bm = (torch.randn(256, 128, device="cuda") * 0.5).half()
d4 = ops.gemm_f16(eng.ctx, A, Bm, None, impl=3)
torch.cuda.synchronize()
print("flow_sums", tuple(flow.shape), int(fsums.sum()), "gemm4cta", bool(torch.isfinite(d4).all()))
img = torch.randint(0, 256, (2, 224, 224, 3), dtype=torch.uint8, device="cuda")
maps = ops.resnet50_maps(eng.ctx, img)
tok = ops.vitb16_tokens(eng.ctx, img)
bgr = ops.yuv420p_to_bgr(torch.randint(0, 256, (2, 48 * 64 * 3 // 2), dtype=torch.uint8, device="cuda"), 48, 64)
torch.cuda.synchronize()
print("maps", len(maps), tuple(tok.shape), tuple(bgr.shape))
tr = HeadTrainer(96, 32, drop_rate=0.1)
X, y = torch.rand(24, 96), torch.rand(24) * 50
for _ in range(2):
    loss = tr.step(X, y, 0.05, 0.9, 0.005, 0.6, 1.0)
tr.swa_update(); tr.update_bn([X]); p = tr.predict(X, swa=True)
torch.cuda.synchronize()
print("trainer", float(loss), bool(torch.isfinite(p).all()))
tr.close(); eng.close()
