"""One mixed-resolution batch through the real sharder on WORLD_SIZE ranks (torchrun): LPT plan -> per-rank predict ->
ragged NCCL all-gather in global video order; rank 0 saves the gathered [V, 35203] matrix and the scores.
Used by tests/test_gpu_multigpu.py (the N-rank result must equal the 1-rank result bit for bit) and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \\
        tools/multigpu_check.py --out gpurun_out/mg2.npz
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relax_vqa_b200 import sharding, synth, weights  # noqa: E402
from relax_vqa_b200.engine import Clip, Engine  # noqa: E402

SPECS = [(272, 480, 3), (144, 256, 2), (360, 640, 2), (100, 150, 1), (272, 480, 1), (333, 517, 2), (144, 256, 3)]   # (H, W, pairs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
    plan = sharding.shard_videos([sharding.video_cost(p, h, w) for h, w, p in SPECS], world)
    clips = []
    for gi in plan[rank]:
        h, w, p = SPECS[gi]
        fr, nx = synth.make_clip(700 + gi, h, w, p)
        clips.append(Clip(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda()))
    if clips:
        feats, score = eng.predict(clips, "konvid_1k")
    else:
        feats = torch.empty((0, 35203), device=eng.device)
        score = torch.empty((0,), device=eng.device)
    feats = sharding.gather_rows(feats, plan)
    score = sharding.gather_rows(score.reshape(-1, 1), plan).reshape(-1)
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(a.out, feats=feats.cpu().numpy(), score=score.cpu().numpy(), plan=np.array([len(p) for p in plan]))
        print("saved", a.out, tuple(feats.shape), [len(p) for p in plan])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
