"""Large-shape sanity run on the GPU: 2160p-20s (43 pairs) and 540p-8s clips through Engine.predict; prints timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relax_vqa_b200 import weights
from relax_vqa_b200.engine import Engine, synthetic_clips_on_device

eng = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
for name, (H, W, pairs, n) in {"540p-8s": (540, 960, 18, 8), "2160p-20s": (2160, 3840, 43, 1), "portrait-720x404": (720, 404, 10, 2)}.items():
    clips = synthetic_clips_on_device(n, H, W, pairs, eng.device, seed=3)
    for _ in range(2):
        feats, score = eng.predict(clips, "konvid_1k")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        feats, score = eng.predict(clips, "konvid_1k")
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {n} clips/step, {dt * 1e3:.1f} ms/step, {n / dt:.1f} videos/s, finite={bool(torch.isfinite(feats).all())}, "
          f"mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB + ws")
    del clips
    torch.cuda.empty_cache()
