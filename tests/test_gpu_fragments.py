"""GPU parity (through the C ABI) for the integer stages A1-A4, A8, A9: bit-exact vs the oracle
and vs the reference's shipped example PNGs."""
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import fragments as F
from oracle import resize as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from relax_vqa_b200 import ops
    c = ops.Context(0)
    yield c
    c.close()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check_pair(frame, nxt, want_exact_positions=True):
    from relax_vqa_b200 import ops
    f, n = _dev(frame[None]), _dev(nxt[None])
    r = ops.absdiff_patchsum(f, n, want_residual=True, want_gray=True)
    res = F.absdiff(nxt, frame)
    assert np.array_equal(r["residual"][0].cpu().numpy(), res)
    sums = F.patch_sums(res)
    assert np.array_equal(r["sums"][0].cpu().numpy().astype(np.float64), sums)
    assert np.array_equal(r["gray0"][0].cpu().numpy(), F.bgr2gray(frame))
    assert np.array_equal(r["gray1"][0].cpu().numpy(), F.bgr2gray(nxt))
    pos, cnt = ops.topk_patches(r["sums"])
    ref_pos = F.topk_positions(sums)
    k = int(cnt[0])
    assert k == len(ref_pos)
    assert [tuple(p) for p in pos[0, :k].cpu().numpy().tolist()] == ref_pos    # stable tie rule == oracle, always
    ori, diff = ops.gather_fragments(f, n, pos, cnt)
    assert np.array_equal(diff[0].cpu().numpy(), F.gather_fragment(res, ref_pos))
    assert np.array_equal(ori[0].cpu().numpy(), F.gather_fragment(frame, ref_pos))
    return ori[0].cpu().numpy(), diff[0].cpu().numpy()


@pytest.mark.parametrize("idx", [2, 3, 4])
def test_shipped_540p_examples(example_dir, idx):
    from relax_vqa_b200 import ops
    d = os.path.join(example_dir, "original_5636101558")
    p = lambda s: cv2.imread(os.path.join(d, f"5636101558_{idx}{s}.png"))
    ori, diff = _check_pair(p(""), p("_next"))
    assert np.array_equal(diff, p("_residual_imp"))
    assert np.array_equal(ori, p("_ori_frag"))
    # flow fragment + merge KATs from the stored flow image
    of = _dev(p("_residual_of")[None])
    pos, cnt = ops.topk_patches(ops.patchsum(of))
    of_frag, _ = ops.gather_fragments(of, None, pos, cnt, want_diff=False)
    assert np.array_equal(of_frag[0].cpu().numpy(), p("_residual_of_imp"))
    merged = ops.merge_fragments(_dev(diff), of_frag[0])
    assert np.array_equal(merged.cpu().numpy(), p("_residual_merged_frag"))


def test_shipped_1080p_and_2160p_examples(example_dir):
    from relax_vqa_b200 import ops
    d = os.path.join(example_dir, "original_TelevisionClip_1080P-68c6")
    p = lambda s: cv2.imread(os.path.join(d, f"TelevisionClip_1080P-68c6_1{s}.png"))
    ori, diff = _check_pair(p(""), p("_next"))
    assert np.array_equal(diff, p("_residual_imp")) and np.array_equal(ori, p("_ori_frag"))
    d = os.path.join(example_dir, "original_Sports_2160P-0455")
    p = lambda s: cv2.imread(os.path.join(d, f"Sports_2160P-0455_1{s}.png"))
    of = _dev(p("_residual_of")[None])
    pos, cnt = ops.topk_patches(ops.patchsum(of))
    of_frag, _ = ops.gather_fragments(of, None, pos, cnt, want_diff=False)
    assert np.array_equal(of_frag[0].cpu().numpy(), p("_residual_of_imp"))
    merged = ops.merge_fragments(_dev(p("_residual_imp")), of_frag[0])
    assert np.array_equal(merged.cpu().numpy(), p("_residual_merged_frag"))


@pytest.mark.parametrize("hw", [(272, 480), (100, 150), (37, 50), (540, 950), (333, 517), (16, 16)])
def test_ragged_sizes_and_small_grids(hw):
    rng = np.random.default_rng(hw[0])
    a = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-20, 20, a.shape), 0, 255).astype(np.uint8)
    _check_pair(a, b)


def test_tie_heavy_and_constant_inputs():
    a = np.zeros((320, 320, 3), np.uint8)
    _check_pair(a, a)                                   # all sums equal: first 196 raster cells
    rng = np.random.default_rng(5)
    b = a.copy()
    b[::16, ::16, 0] = rng.integers(0, 3, (20, 20))       # only 3 distinct sums -> massive ties
    _check_pair(a, b)


def test_batched_matches_single():
    from relax_vqa_b200 import ops, synth
    fr, nx = synth.make_clip(11, 272, 480, 4)
    r = ops.absdiff_patchsum(_dev(fr), _dev(nx))
    pos, cnt = ops.topk_patches(r["sums"])
    ori, diff = ops.gather_fragments(_dev(fr), _dev(nx), pos, cnt)
    for t in range(4):
        res = F.absdiff(nx[t], fr[t])
        frag, p, s = F.process_patches(res)
        assert np.array_equal(diff[t].cpu().numpy(), frag)
        assert np.array_equal(ori[t].cpu().numpy(), F.gather_fragment(fr[t], p))


def test_merge_round_half_even_exhaustive():
    from relax_vqa_b200 import ops
    a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    a, b = a.reshape(-1)[:65535].copy(), b.reshape(-1)[:65535].copy()     # odd length: exercises the tail path
    out = ops.merge_fragments(_dev(a), _dev(b)).cpu().numpy()
    assert np.array_equal(out, F.merge_fragments(a, b))
    assert np.array_equal(out, cv2.addWeighted(a, 0.5, b, 0.5, 0).reshape(-1))


@pytest.mark.parametrize("wh", [(960, 540), (1920, 1080), (404, 720), (224, 224), (300, 224), (224, 150), (97, 61), (3840, 2160)])
@pytest.mark.parametrize("filt", [0, 1])
def test_resize_bit_exact(ctx, wh, filt):
    from relax_vqa_b200 import ops
    from PIL import Image
    rng = np.random.default_rng(wh[0] + filt)
    img = rng.integers(0, 256, (2, wh[1], wh[0], 3), dtype=np.uint8)
    out = ops.resize_pil(ctx, _dev(img), filt).cpu().numpy()
    pil = Image.BILINEAR if filt == 0 else Image.LANCZOS
    for i in range(2):
        ref = np.asarray(Image.fromarray(img[i]).resize((224, 224), pil))
        assert np.array_equal(out[i], ref)
        assert np.array_equal(out[i], R.resize(img[i], 224, 224, filt))
    sw = ops.resize_pil(ctx, _dev(img), filt, swap_rb=True).cpu().numpy()
    assert np.array_equal(sw, out[..., ::-1])


@pytest.mark.parametrize("wh", [(960, 540), (1920, 1080), (404, 720), (224, 300), (97, 61), (3840, 2160)])
def test_resize_pair_equals_two_calls(ctx, wh):
    """Both filters from one staged read of the source (the engine's call): bit-identical to the two single-filter calls,
    which test_resize_bit_exact pins to Pillow.  Includes saturated images (extreme accumulator values of the split sums)."""
    from relax_vqa_b200 import ops
    rng = np.random.default_rng(wh[0])
    img = rng.integers(0, 256, (3, wh[1], wh[0], 3), dtype=np.uint8)
    img[1] = 255
    img[2] = np.where(rng.random(img[2].shape) < 0.5, 0, 255)
    for swap in (False, True):
        bil, lan = ops.resize_pil_pair(ctx, _dev(img), swap_rb=swap)
        assert torch.equal(bil, ops.resize_pil(ctx, _dev(img), 0, swap_rb=swap))
        assert torch.equal(lan, ops.resize_pil(ctx, _dev(img), 1, swap_rb=swap))
