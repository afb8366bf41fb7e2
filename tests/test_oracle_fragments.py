"""Pins oracle/fragments.py (A1-A4, A7, A8) against the reference's shipped example PNGs
(visualisation/visualisation_example) and the reference-generated
fixtures of tests/golden/gen_golden.py.  Bit-exact."""
import os

import cv2
import numpy as np
import pytest

from oracle import fragments as F


def _rd(path):
    img = cv2.imread(path)
    assert img is not None, path
    return img


@pytest.mark.parametrize("idx", [2, 3, 4])
def test_540p_example_chain(example_dir, idx):
    d = os.path.join(example_dir, "original_5636101558")
    p = lambda s: os.path.join(d, f"5636101558_{idx}{s}.png")
    a, b = _rd(p("")), _rd(p("_next"))
    res = F.absdiff(b, a)
    assert np.array_equal(res, _rd(p("_residual")))
    frag, pos, sums = F.process_patches(res)
    assert sums.shape == (33, 60) and not F.cut_is_tied(sums)
    assert np.array_equal(frag, _rd(p("_residual_imp")))
    assert np.array_equal(F.gather_fragment(a, pos), _rd(p("_ori_frag")))
    of_frag, _, _ = F.process_patches(_rd(p("_residual_of")))
    assert np.array_equal(of_frag, _rd(p("_residual_of_imp")))
    assert np.array_equal(F.merge_fragments(_rd(p("_residual_imp")), _rd(p("_residual_of_imp"))),
                          _rd(p("_residual_merged_frag")))


def test_1080p_example_chain(example_dir):
    d = os.path.join(example_dir, "original_TelevisionClip_1080P-68c6")
    p = lambda s: os.path.join(d, f"TelevisionClip_1080P-68c6_1{s}.png")
    a, b = _rd(p("")), _rd(p("_next"))
    res = F.absdiff(b, a)
    frag, pos, sums = F.process_patches(res)
    assert sums.shape == (67, 120) and not F.cut_is_tied(sums)
    assert np.array_equal(frag, _rd(p("_residual_imp")))
    assert np.array_equal(F.gather_fragment(a, pos), _rd(p("_ori_frag")))
    of_frag, _, _ = F.process_patches(_rd(p("_residual_of")))
    assert np.array_equal(of_frag, _rd(p("_residual_of_imp")))
    assert np.array_equal(F.merge_fragments(frag, of_frag), _rd(p("_residual_merged_frag")))


def test_2160p_example_downstream(example_dir):
    d = os.path.join(example_dir, "original_Sports_2160P-0455")
    p = lambda s: os.path.join(d, f"Sports_2160P-0455_1{s}.png")
    of_frag, _, sums = F.process_patches(_rd(p("_residual_of")))
    assert sums.shape == (135, 240)
    assert np.array_equal(of_frag, _rd(p("_residual_of_imp")))
    assert np.array_equal(F.merge_fragments(_rd(p("_residual_imp")), of_frag), _rd(p("_residual_merged_frag")))


def test_reference_generated_fixture(golden_dir):
    from relax_vqa_b200 import synth
    g = np.load(os.path.join(golden_dir, "ref_fragments_synth.npz"))
    fr, nx = synth.make_clip(int(g["clip_seed"]), int(g["H"]), int(g["W"]), int(g["T"]))
    for t in range(int(g["T"])):
        res = F.absdiff(nx[t], fr[t])
        frag, pos, sums = F.process_patches(res)
        assert np.array_equal(sums, g[f"sums{t}"])
        if not F.cut_is_tied(sums):
            assert np.array_equal(np.array(pos), g[f"pos{t}"])
            assert np.array_equal(frag, g[f"diff_frag{t}"])
            assert np.array_equal(F.gather_fragment(fr[t], pos), g[f"ori_frag{t}"])
        ffrag, fpos, fsums = F.process_patches(g[f"flow_rgb{t}"])
        if not F.cut_is_tied(fsums):
            assert np.array_equal(np.array(fpos), g[f"flow_pos{t}"])
            assert np.array_equal(ffrag, g[f"flow_frag{t}"])
        assert np.array_equal(F.merge_fragments(g[f"diff_frag{t}"], g[f"flow_frag{t}"]), g[f"merged{t}"])


def test_edge_cases():
    rng = np.random.default_rng(0)
    # fewer than 196 patches: canvas is zero-filled past the last patch
    small = rng.integers(0, 256, (100, 150, 3), dtype=np.uint8)       # 6 x 9 = 54 patches
    frag, pos, sums = F.process_patches(small)
    assert len(pos) == 54 and pos == sorted(pos)
    assert np.array_equal(frag[48:64, 14 * 16 - 32:], np.zeros((16, 32, 3), np.uint8)[:, :32])
    assert frag[64:].sum() == 0
    # all-equal sums: stable tie rule keeps the first 196 raster positions
    flat = np.zeros((320, 320, 3), np.uint8)
    _, pos, sums = F.process_patches(flat)
    assert F.cut_is_tied(sums) and pos == [(i // 20, i % 20) for i in range(196)]
    # ragged size: crop to a multiple of 16
    assert F.patch_sums(np.ones((37, 50, 3), np.uint8)).shape == (2, 3)
    # round-half-even merge
    a = np.array([[[1, 2, 3]], [[255, 0, 7]]], np.uint8)
    b = np.array([[[2, 3, 4]], [[254, 1, 7]]], np.uint8)
    assert F.merge_fragments(a, b).tolist() == [[[2, 2, 4]], [[254, 0, 7]]]
    assert np.array_equal(F.merge_fragments(a, b), cv2.addWeighted(a, 0.5, b, 0.5, 0))
    x = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    y = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    assert np.array_equal(F.merge_fragments(x, y), cv2.addWeighted(x, 0.5, y, 0.5, 0))
    assert np.array_equal(F.absdiff(x, y), cv2.absdiff(x, y))
    assert np.array_equal(F.bgr2gray(x), cv2.cvtColor(x, cv2.COLOR_BGR2GRAY))
