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


def test_collation_matches_the_unmodified_reference_script(tmp_path, golden_dir):
    """tests/golden/ref_formats.npz = the reference's extract_npy2mat.py run as a script on .npy files written by this repo's
    writer: same matrix (float64, temporal mean, bit for bit), same .mat file name and key."""
    import scipy.io
    g = np.load(os.path.join(golden_dir, "ref_formats.npz"))
    d = str(tmp_path / "feat")
    mats = [g[f"in{i}"] for i in range(3)]
    for i, m in enumerate(mats):
        fmt.save_video_npy(d, i, "resnet50", m)
    out = fmt.collate(d, 3, "resnet50")
    assert out.dtype == g["matrix"].dtype and np.array_equal(out, g["matrix"])
    p = fmt.save_features(str(tmp_path / "pool" / "original_features"), "konvid_1k", out, "resnet50")
    assert os.path.basename(p) == os.path.basename(str(g["rel_path"]))
    assert np.array_equal(scipy.io.loadmat(p)[str(g["key"])], g["matrix"])


def test_reference_script_runs_on_our_files_live():
    """Same check against the reference tree itself when it is present (build container only)."""
    import pytest
    import sys
    if not os.path.isdir("/root/reference/src/data_processing"):
        pytest.skip("reference tree not present (GPU box)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import gen_golden_formats as G
    mats, vids = G.inputs()
    matrix, rel, key = G.run_reference_script(mats, vids)
    assert key == "konvid_1k" and matrix.shape == (3, 11)
    assert np.array_equal(matrix, np.stack([m.mean(axis=0) for m in mats]).astype(np.float64))
