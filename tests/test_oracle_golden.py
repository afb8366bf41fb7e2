"""Pins oracle/backbones.py, head.py, pipeline.py (A10-A18) against outputs of the UNMODIFIED
reference run with the same seeded weights (tests/golden/gen_golden.py)."""
import os

import numpy as np
import pytest

from oracle import pipeline as P
from relax_vqa_b200 import synth, weights


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_video_synth.npz"))


def test_video_blocks_vector_and_score(golden, golden_dir):
    g = golden
    fr, nx = synth.make_clip(int(g["clip_seed"]), int(g["H"]), int(g["W"]), int(g["T"]))
    rsd = weights.seeded_resnet50_state_dict(int(g["resnet_seed"]))
    vsd = weights.seeded_vitb16_state_dict(int(g["vit_seed"]))
    import cv2
    # strict leg: the reference's own flow (cv2) feeds the restated pipeline
    blocks = P.video_feature_blocks(fr, nx, rsd, vsd,
                                    flow_fn=lambda a, b: cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0))
    for k, width in (("full_resnet", 13120), ("full_vit", 2304), ("frag_resnet", 15171), ("frag_vit", 4608)):
        assert blocks[k].shape == (int(g["T"]), width)
        # fp32 CPU vs fp32 CPU, different conv batching: 1e-4 of the block's RMS
        rms = np.sqrt(np.mean(g[k] ** 2))
        assert np.abs(blocks[k] - g[k]).max() <= 1e-4 * rms, k
    vec = P.video_vector(blocks)
    assert vec.shape == (35203,)
    assert np.abs(vec - g["vector"]).max() < 1e-4
    # full-oracle leg: restated Farneback; a few +-1 flow-colour pixels may move the merged fragment
    blocks2 = P.video_feature_blocks(fr[:1], nx[:1], rsd, vsd)
    for k in ("frag_resnet", "frag_vit"):
        rms = np.sqrt(np.mean(g[k] ** 2))
        assert np.abs(blocks2[k] - g[k][:1]).max() <= 2e-3 * rms, k
    s = np.load(os.path.join(golden_dir, "konvid_1k_scaler_imputer.npz"))
    hsd = weights.fix_state_dict(weights.seeded_head_state_dict(int(g["head_seed"]), swa_format=True))
    score = P.predict(vec, hsd, s["imputer_mean"], s["scale"], s["minv"], "konvid_1k")
    assert abs(score - float(g["score"])) < 1e-4


def test_shipped_scaler_is_identity_and_imputer_fills(golden_dir):
    from oracle import head as HD
    s = np.load(os.path.join(golden_dir, "konvid_1k_scaler_imputer.npz"))
    assert s["scale"].shape == (35203,) and np.abs(s["scale"] - 1).max() < 1e-14 and np.all(s["minv"] == 0)
    x = np.zeros((1, 35203))
    x[0, 5] = np.nan
    y = HD.impute_scale(x, s["imputer_mean"], s["scale"], s["minv"])
    assert y[0, 5] == s["imputer_mean"][5] and not np.isnan(y).any()


def test_state_dict_formats():
    sd = weights.seeded_head_state_dict(1, swa_format=True)
    assert "n_averaged" in sd and all(k == "n_averaged" or k.startswith("module.") for k in sd)
    fixed = weights.fix_state_dict(sd)
    assert [k for k, _ in weights.head_spec()] == list(fixed.keys())
    assert fixed["fc1.weight"].shape == (256, 35203)
    import torchvision
    m = torchvision.models.resnet50(weights=None)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in weights.resnet50_spec()]
