"""Generates tests/golden/ref_formats.npz: the UNMODIFIED reference script src/data_processing/extract_npy2mat.py run (as
__main__, with its hard-coded settings: konvid_1k / resnet50 / pool / frame_diff_frag) on per-video .npy files written by THIS
repo's writer (relax_vqa_b200.data_processing.extract_npy2mat.save_video_npy) - SURVEY.md 8(f) row 2.

Run in the build container only (needs /root/reference):   python tests/golden/gen_golden_formats.py
"""
import os
import runpy
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RELAXVQA_REF", "/root/reference")
sys.path.insert(0, ROOT)
from relax_vqa_b200.data_processing import extract_npy2mat as fmt  # noqa: E402


def run_reference_script(mats, vids):
    """-> (matrix, mat file path relative to the scratch root, key) produced by the reference script on our files."""
    import pandas as pd
    import scipy.io
    scratch = tempfile.mkdtemp(prefix="relaxvqa_formats_")
    cwd = os.path.join(scratch, "src", "data_processing")
    os.makedirs(cwd)
    os.makedirs(os.path.join(scratch, "metadata"))
    pd.DataFrame(dict(vid=vids, mos=[3.0] * len(vids))).to_csv(os.path.join(scratch, "metadata", "KONVID_1K_metadata.csv"), index=False)
    feat_dir = fmt.features_dir(os.path.join(scratch, "features_residual_frag"), "resnet50", "pool", "konvid_1k")
    for i, m in enumerate(mats):
        fmt.save_video_npy(feat_dir, i, "resnet50", m)                      # OUR writer, the reference's naming
    old = os.getcwd()
    os.chdir(cwd)
    try:
        runpy.run_path(os.path.join(REF, "src", "data_processing", "extract_npy2mat.py"), run_name="__main__")
    finally:
        os.chdir(old)
    rel = os.path.join("features_residual_frag", "pool", "original_features", "resnet50_konvid_1k_original_features.mat")
    mat = scipy.io.loadmat(os.path.join(scratch, rel))
    keys = [k for k in mat if not k.startswith("__")]
    return mat[keys[0]], rel, keys[0]


def inputs():
    rng = np.random.default_rng(3)
    return [rng.standard_normal((t, 11)).astype(np.float32) for t in (4, 1, 7)], ["vidA", "vidB", "vidC"]


if __name__ == "__main__":
    mats, vids = inputs()
    matrix, rel, key = run_reference_script(mats, vids)
    np.savez_compressed(os.path.join(HERE, "ref_formats.npz"), matrix=matrix, rel_path=rel, key=key,
                        **{f"in{i}": m for i, m in enumerate(mats)})
    print("reference script output", matrix.shape, matrix.dtype, rel, key)
