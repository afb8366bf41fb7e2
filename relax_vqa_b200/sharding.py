"""Video sharding across the GPUs of one box and the end-of-step gather.

Videos are independent until the head, so ranks never exchange data on the path; the only collective
is an all-gather of the per-video feature rows and scores (NCCL on GPUs, gloo in the CPU tests)."""
from typing import List, Sequence

import torch
import torch.distributed as dist

PAIR_BYTES_PER_PIXEL = 335.0          # bandwidth stages, SURVEY.md 8(d)
DENSE_FLOP_PER_PAIR = 129.9e9


def video_cost(pairs: int, height: int, width: int, hbm_gbs=6461.5, tflops=1381.9) -> float:
    """Roofline cost model (seconds) used to balance mixed-resolution batches."""
    return pairs * (PAIR_BYTES_PER_PIXEL * height * width / (hbm_gbs * 1e9) + DENSE_FLOP_PER_PAIR / (tflops * 1e12))


def shard_videos(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment; returns, per rank, the sorted video indices it owns.
    Deterministic (ties broken by index) so every rank computes the same plan without communication."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owned: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owned[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in owned]


def gather_rows(local: torch.Tensor, owned: List[List[int]], group=None) -> torch.Tensor:
    """All-gather ragged per-rank row blocks and return them in global video order on every rank.
    local: [len(owned[rank]), D]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    n_max = max(len(o) for o in owned)
    D = local.shape[1:]
    pad = torch.zeros((n_max,) + tuple(D), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    total = sum(len(o) for o in owned)
    out = torch.empty((total,) + tuple(D), dtype=local.dtype, device=local.device)
    for r, idx in enumerate(owned):
        if idx:
            out[torch.tensor(idx, device=local.device)] = bufs[r][:len(idx)]
    return out
