"""Host logic of the multi-GPU path on CPU: LPT sharding + ragged all-gather over gloo, world_size 2."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relax_vqa_b200 import sharding


def test_lpt_sharding_is_balanced_and_deterministic():
    costs = [sharding.video_cost(22, 1080, 1920)] * 5 + [sharding.video_cost(18, 540, 960)] * 7 + [sharding.video_cost(43, 2160, 3840)]
    plan = sharding.shard_videos(costs, 4)
    assert sorted(i for p in plan for i in p) == list(range(len(costs)))
    assert plan == sharding.shard_videos(costs, 4)
    loads = [sum(costs[i] for i in p) for p in plan]
    assert max(loads) <= max(costs) + sum(costs) / 4       # LPT bound
    assert sharding.shard_videos(costs, 1) == [list(range(len(costs)))]
    assert sharding.video_cost(22, 1080, 1920) > sharding.video_cost(22, 540, 960)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [3.0, 1.0, 1.0, 2.0, 1.0]                      # ragged: rank 0 gets 2 videos, rank 1 gets 3
    plan = sharding.shard_videos(costs, world)
    mine = plan[rank]
    feats = torch.stack([torch.full((7,), float(i)) for i in mine])
    scores = torch.tensor([[10.0 * i] for i in mine])
    all_f = sharding.gather_rows(feats, plan)
    all_s = sharding.gather_rows(scores, plan)
    ok = torch.equal(all_f[:, 0], torch.arange(5.0)) and torch.equal(all_s[:, 0], 10.0 * torch.arange(5.0))
    ret[rank] = bool(ok) and len(plan[0]) != len(plan[1])
    dist.destroy_process_group()


def test_ragged_gather_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
