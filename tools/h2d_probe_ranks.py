"""Attributes the end-to-end (host buffers) scaling loss at N ranks: per-rank H2D rate when every rank
copies at once vs when one rank copies alone, compute-only step time, pipelined host-to-score step time.  Launch with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 tools/h2d_probe_ranks.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relax_vqa_b200 import weights  # noqa: E402
from relax_vqa_b200.engine import Engine, bind_host_to_gpu, synthetic_clips_on_device  # noqa: E402

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
cpus = bind_host_to_gpu(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(local, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
clips = synthetic_clips_on_device(4, 1080, 1920, 22, eng.device, seed=1000 + rank)
host = [(c.frames.cpu().pin_memory(), c.nexts.cpu().pin_memory()) for c in clips]
nbytes = sum(f.numel() + n.numel() for f, n in host)
STEPS = 8


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, n=STEPS):
    fn(3)
    barrier()
    t = time.perf_counter()
    fn(n)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


def copies(n):
    for _ in range(n):
        d = [(f.to(eng.device, non_blocking=True), x.to(eng.device, non_blocking=True)) for f, x in host]
        del d


def compute(n):
    for _ in range(n):
        eng.predict(clips, "live_vqc")


def pipe_loop(n):
    t = eng.submit_host(host, "live_vqc")
    for i in range(n):
        nx = eng.submit_host(host, "live_vqc") if i + 1 < n else None
        eng.result(t)
        t = nx


res = dict(rank=rank, cpus_bound=len(cpus) if cpus else None)
res["h2d_all_ranks_gbs"] = nbytes / timed(copies) / 1e6
# one rank at a time: the others wait at the barrier
solo = None
for r in range(world):
    if r == rank:
        copies(2)
        torch.cuda.synchronize()
        t = time.perf_counter()
        copies(STEPS)
        torch.cuda.synchronize()
        solo = nbytes / ((time.perf_counter() - t) / STEPS * 1e3) / 1e6
    barrier()
res["h2d_alone_gbs"] = solo
res["compute_ms_per_step"] = timed(compute)
res["e2e_pipelined_ms_per_step"] = timed(pipe_loop)
if world > 1:
    parts = [None] * world
    dist.all_gather_object(parts, res)
else:
    parts = [res]
if rank == 0:
    agg = dict(world=world, bytes_per_step_per_rank=nbytes, ranks=parts,
               sum_h2d_all_ranks_gbs=sum(p["h2d_all_ranks_gbs"] for p in parts),
               mean_h2d_alone_gbs=sum(p["h2d_alone_gbs"] for p in parts) / world,
               needed_gbs_per_rank_to_hide_copy=nbytes / (sum(p["compute_ms_per_step"] for p in parts) / world) / 1e6)
    print(json.dumps(agg))
if world > 1:
    dist.destroy_process_group()
