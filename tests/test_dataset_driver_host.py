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
