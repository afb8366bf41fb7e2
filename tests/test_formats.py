"""CPU checks of the on-disk formats against the reference's conventions."""
import os

import numpy as np

from relax_vqa_b200.data_processing import extract_npy2mat as fmt


def test_names_and_collation(tmp_path):
    assert fmt.npy_name(0, "resnet50") == "video_1_resnet50_feature_map_original.npy"         # main_fragment_layerstack.py:353
    assert fmt.npy_name(4, "vit", resolution="360P") == "video_5_vit_feature_map_original_360P.npy"
    assert fmt.features_dir("../features", "vit", "pool", "konvid_1k") == "../features/vit/pool/konvid_1k/original/"
    d = str(tmp_path / "f")
    rng = np.random.default_rng(0)
    mats = [rng.standard_normal((t, 7)).astype(np.float32) for t in (3, 5)]
    for i, m in enumerate(mats):
        fmt.save_video_npy(d, i, "resnet50", m)
    out = fmt.collate(d, 2, "resnet50")
    assert out.shape == (2, 7) and out.dtype == np.float64
    assert np.array_equal(out[0], np.mean(mats[0], axis=0)) and np.array_equal(out[1], np.mean(mats[1], axis=0))
    import scipy.io
    p = fmt.save_features(str(tmp_path / "m"), "live_vqc", out, "resnet50")
    assert os.path.basename(p) == "resnet50_live_vqc_original_features.mat"
    assert np.array_equal(scipy.io.loadmat(p)["live_vqc"], out)
