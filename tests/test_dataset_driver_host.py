"""Host logic of the dataset-scale job that needs no GPU: output naming, atomic / validated
resume, greyscale filtering, the cost-balanced plan two ranks compute independently."""
import os

import numpy as np
import pytest

from relax_vqa_b200 import dataset_driver as dd
from relax_vqa_b200 import sharding


def test_output_paths_follow_the_reference_layout(tmp_path):
    p = dd.output_paths("out", "konvid_1k", 4)
    assert p["frag_resnet"] == "out/features_merged_frag/resnet50/layer_stack/konvid_1k/original/video_5_resnet50_feature_map_original.npy"
    assert p["full_vit"] == "out/features/vit/pool/konvid_1k/original/video_5_vit_feature_map_original.npy"
    with pytest.raises(ValueError):
        dd.output_paths("out", "youtube_ugc", 0)                            # no '.../original_None/' folders
    u = dd.output_paths("out", "youtube_ugc", 0, resolution="360P")
    assert u["full_resnet"] == "out/features/resnet50/layer_stack/resolution_ugc/original_360P/video_1_resnet50_feature_map_original_360P.npy"


def test_atomic_save_and_validated_resume(tmp_path):
    path = str(tmp_path / "a" / "video_1_vit_feature_map_original.npy")
    m = np.arange(12, dtype=np.float32).reshape(3, 4)
    dd.atomic_save_npy(path, m)
    assert np.array_equal(np.load(path), m) and os.listdir(os.path.dirname(path)) == [os.path.basename(path)]   # no temp file left
    assert dd._is_complete(path)
    with open(path, "r+b") as f:                                            # a writer killed mid-file
        f.truncate(os.path.getsize(path) - 16)
    assert not dd._is_complete(path) and not dd._is_complete(str(tmp_path / "missing.npy"))


def test_greyscale_report_and_pair_count(tmp_path):
    import pandas as pd
    csv = tmp_path / "YOUTUBE_UGC_greyscale_metadata.csv"
    pd.DataFrame({"index": [3, 17], "vid": ["a", "b"]}).to_csv(csv, index=False)
    assert dd.greyscale_indices(str(csv)) == {3, 17}                         # first column, split_train_test.py:113-117
    assert dd.sampled_pairs(300, 29.97) == 22 and dd.sampled_pairs(240, 29.97) == 18 and dd.sampled_pairs(600, 29.97) == 43
    assert dd.sampled_pairs(15, 29.97) == 1 and dd.sampled_pairs(1, 29.97) == 0 and dd.sampled_pairs(100, 90000) == 1


def test_ranks_compute_the_same_balanced_plan():
    rng = np.random.default_rng(0)
    rows = [(int(rng.integers(8, 40)), *[(720, 1280), (480, 640), (1080, 1920), (360, 640)][int(rng.integers(0, 4))]) for _ in range(200)]
    costs = [sharding.video_cost(p, h, w) for p, h, w in rows]
    plans = [sharding.shard_videos(costs, 8) for _ in range(2)]             # every rank runs this independently
    assert plans[0] == plans[1]
    loads = [sum(costs[i] for i in r) for r in plans[0]]
    assert max(loads) / (sum(loads) / 8) < 1.02                             # finish times within 2 %


# ---- the whole job on two ranks (gloo, CPU) with a stand-in engine: sharding, per-rank extraction, status gather, resume -------
class _FakeEngine:
    """Duck-typed Engine for the host logic: features are simple functions of the frames, so any rank computes the same
    rows for the same video and the test can check who wrote what."""

    def __init__(self):
        import torch
        self.device = torch.device("cpu")
        self.calls = []

    def extract_blocks(self, clips):
        import torch
        self.calls.append(len(clips))
        fo, po = [0], [0]
        for c in clips:
            fo.append(fo[-1] + c.frames.shape[0]); po.append(po[-1] + c.nexts.shape[0])
        fr = torch.cat([c.frames.float().mean(dim=(1, 2, 3)) for c in clips])        # clips of one batch may differ in resolution
        nx = torch.cat([c.nexts.float().mean(dim=(1, 2, 3)) for c in clips])
        wide = lambda t, w: t[:, None].expand(-1, w).contiguous()
        return dict(full_resnet=wide(fr, 13120), full_vit=wide(fr, 2304), frag_stack=wide(nx, 13120), frag_pool=wide(nx, 2051),
                    frag_vit_ori=wide(nx, 2304), frag_vit_mer=wide(nx, 2304), full_off=torch.tensor(fo), pair_off=torch.tensor(po))


def _write_job(root, n=5):
    import cv2
    import pandas as pd
    rng = np.random.default_rng(0)
    sizes = [(32, 48), (16, 32), (32, 48), (48, 64), (16, 32)][:n]
    for i, (h, w) in enumerate(sizes):
        d = os.path.join(root, "frames", f"video_{i + 1}")
        os.makedirs(d)
        for k in range(2):
            cv2.imwrite(os.path.join(d, f"v{i}_{k + 1}.png"), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
            cv2.imwrite(os.path.join(d, f"v{i}_{k + 1}_next.png"), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    csv = os.path.join(root, "meta.csv")
    pd.DataFrame(dict(vid=[f"v{i}" for i in range(n)], width=[s[1] for s in sizes], height=[s[0] for s in sizes], framerate=[29.97] * n,
                      nb_frames=[30, 60, 30, 90, 60][:n])).to_csv(csv, index=False)
    grey = os.path.join(root, "grey.csv")
    pd.DataFrame({"index": [2], "vid": ["v2"]}).to_csv(grey, index=False)
    return csv, grey


def _job_worker(rank, world, port, root, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine()
    csv, grey = os.path.join(root, "meta.csv"), os.path.join(root, "grey.csv")
    st = dd.run(csv, os.path.join(root, "frames"), os.path.join(root, "out"), "konvid_1k", engine=eng, batch_videos=2, greyscale_csv=grey)
    ret[rank] = (st, sum(eng.calls))
    dist.destroy_process_group()


def test_job_on_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    root = str(tmp_path)
    _write_job(root)
    ret = mp.Manager().dict()
    mp.spawn(_job_worker, args=(2, 29700 + os.getpid() % 200, root, ret), nprocs=2, join=True)
    st0, n0 = ret[0]
    st1, n1 = ret[1]
    assert st0 == st1 == [("v0", "done"), ("v1", "done"), ("v2", "greyscale"), ("v3", "done"), ("v4", "done")]    # gathered, metadata order
    assert n0 + n1 == 4 and n0 > 0 and n1 > 0                                         # every video extracted once, by one rank
    p = dd.output_paths(os.path.join(root, "out"), "konvid_1k", 3)
    assert np.load(p["frag_resnet"]).shape == (2, 15171) and not os.path.exists(dd.output_paths(os.path.join(root, "out"), "konvid_1k", 2)["full_vit"])
    # resume on one rank: everything is skipped, a truncated file is redone
    with open(p["full_vit"], "r+b") as f:
        f.truncate(100)
    eng = _FakeEngine()
    st = dd.run(os.path.join(root, "meta.csv"), os.path.join(root, "frames"), os.path.join(root, "out"), "konvid_1k", engine=eng,
                greyscale_csv=os.path.join(root, "grey.csv"), rank=0, world=1)
    assert [s for _, s in st] == ["skipped", "skipped", "greyscale", "done", "skipped"] and sum(eng.calls) == 1
