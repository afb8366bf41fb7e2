"""GPU parity for A10-A18 against the fp32 CPU oracle with the same seeded weights.

Tolerance metric: per feature segment, max|a-b| / rms(b) <= 1e-2 (north_star's
bound for stacked features); the SIMT check path (fp16 operands, fp32 accumulate, same math) must agree
with the tcgen05 path to accumulation-order noise."""
import os

import numpy as np
import pytest
import torch

from oracle import backbones as OB
from oracle import head as OH
from relax_vqa_b200 import synth, weights

pytestmark = pytest.mark.gpu

SEGMENTS = [64, 256, 256, 256, 512, 512, 512, 512, 1024, 1024, 1024, 1024, 2048, 2048, 2048]


@pytest.fixture(scope="module")
def ctx():
    from relax_vqa_b200 import ops
    c = ops.Context(0)
    ops.load_resnet50(c, weights.seeded_resnet50_state_dict(1234))
    ops.load_vitb16(c, weights.seeded_vitb16_state_dict(4321))
    yield c
    c.close()


@pytest.fixture(scope="module")
def images():
    fr, _ = synth.make_clip(21, 224, 224, 5)
    rng = np.random.default_rng(0)
    extra = rng.integers(0, 256, (2, 224, 224, 3), dtype=np.uint8)        # white-noise images too
    return np.concatenate([fr, extra])                                     # BGR uint8


def seg_err(a, b, widths):
    out, o = [], 0
    for w in widths:
        ref = b[:, o:o + w]
        out.append(np.abs(a[:, o:o + w] - ref).max() / max(np.sqrt(np.mean(ref ** 2)), 1e-12))
        o += w
    return out


def test_resnet50_layerstack_and_pool(ctx, images):
    from relax_vqa_b200 import ops
    rsd = weights.seeded_resnet50_state_dict(1234)
    x = OB.resnet_preprocess(images[..., ::-1])
    ref_stack = OB.resnet50_layerstack(rsd, x)
    ref_pool = OB.resnet50_pool(rsd, x)
    dev = torch.from_numpy(images).cuda()
    ctx.set_gemm_impl(1)
    chk_stack, chk_pool = ops.resnet50_features(ctx, dev, is_bgr=True, want_stack=True, want_pool=True)
    ctx.set_gemm_impl(0)
    stack, pool = ops.resnet50_features(ctx, dev, is_bgr=True, want_stack=True, want_pool=True)
    torch.cuda.synchronize()
    stack, pool, chk_stack, chk_pool = (t.cpu().numpy() for t in (stack, pool, chk_stack, chk_pool))
    e_chk = seg_err(chk_stack, ref_stack, SEGMENTS)
    e = seg_err(stack, ref_stack, SEGMENTS)
    print("resnet seg err (simt check):", np.round(e_chk, 5))
    print("resnet seg err (tcgen05):   ", np.round(e, 5))
    assert max(seg_err(stack, chk_stack, SEGMENTS)) < 2e-3          # tensor-core path == check path
    assert max(e) <= 1e-2
    assert max(seg_err(pool, ref_pool, [2048, 1, 1, 1])) <= 1e-2
    assert np.abs(pool[:, :2048] - stack[:, -2048:]).max() == 0.0  # avgpool == last hook


def test_resnet_batch_invariance_and_rgb_flag(ctx, images):
    from relax_vqa_b200 import ops
    dev = torch.from_numpy(images).cuda()
    full, _ = ops.resnet50_features(ctx, dev)
    one, _ = ops.resnet50_features(ctx, dev[3:4].contiguous())
    assert torch.equal(full[3:4], one)                               # bit-identical regardless of batch
    rgb = torch.from_numpy(np.ascontiguousarray(images[..., ::-1])).cuda()
    swapped, _ = ops.resnet50_features(ctx, rgb, is_bgr=False)
    assert torch.equal(full, swapped)


def test_vitb16_pool(ctx, images):
    from relax_vqa_b200 import ops
    vsd = weights.seeded_vitb16_state_dict(4321)
    ref = OB.vit_pool(vsd, OB.vit_preprocess(images[..., ::-1]))
    dev = torch.from_numpy(images).cuda()
    ctx.set_gemm_impl(1)
    chk = ops.vitb16_features(ctx, dev).cpu().numpy()
    ctx.set_gemm_impl(0)
    got = ops.vitb16_features(ctx, dev).cpu().numpy()
    e_chk, e = seg_err(chk, ref, [768, 768, 768]), seg_err(got, ref, [768, 768, 768])
    print("vit seg err (simt check):", np.round(e_chk, 5), " (tcgen05):", np.round(e, 5))
    assert max(seg_err(got, chk, [768, 768, 768])) < 3e-3
    assert max(e) <= 1e-2
    one = ops.vitb16_features(ctx, dev[2:3].contiguous()).cpu().numpy()
    assert np.array_equal(one, got[2:3])


def test_head_and_temporal_mean(ctx, golden_dir):
    from relax_vqa_b200 import ops
    s = np.load(os.path.join(golden_dir, "konvid_1k_scaler_imputer.npz"))
    hsd = weights.seeded_head_state_dict(99, swa_format=True)
    ops.load_head(ctx, hsd, s["imputer_mean"], s["scale"], s["minv"])
    rng = np.random.default_rng(3)
    T = [3, 5, 1]
    blocks = [rng.standard_normal((sum(T), w)).astype(np.float32) for w in (13120, 2304, 13120, 2051, 2304, 2304)]
    off = np.concatenate([[0], np.cumsum(T)]).astype(np.int32)
    d = [torch.from_numpy(b).cuda() for b in blocks]
    o = torch.from_numpy(off).cuda()
    feats = ops.temporal_mean_concat(*d, o, o).cpu().numpy()
    for v in range(3):
        ref = np.concatenate([np.mean(b[off[v]:off[v + 1]], axis=0) for b in blocks])
        assert np.abs(feats[v] - ref).max() < 1e-6
    feats[1, 7] = np.nan                                              # imputer path
    score = ops.head_forward(ctx, torch.from_numpy(feats).cuda()).cpu().numpy()
    x = OH.impute_scale(feats, s["imputer_mean"], s["scale"], s["minv"]).astype(np.float32)
    ref = OH.mlp_forward(weights.fix_state_dict(hsd), x)
    assert np.abs(score - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())


def test_reference_golden_video_blocks(ctx, golden_dir):
    """Per-frame feature matrices of the UNMODIFIED reference (tests/golden/gen_golden.py) on the
    full-frame path: resize (bit-exact) -> backbones."""
    from relax_vqa_b200 import ops
    g = np.load(os.path.join(golden_dir, "ref_video_synth.npz"))
    fr, nx = synth.make_clip(int(g["clip_seed"]), int(g["H"]), int(g["W"]), int(g["T"]))
    dev = torch.from_numpy(fr).cuda()
    rn_in = ops.resize_pil(ctx, dev, ops.BILINEAR)
    vt_in = ops.resize_pil(ctx, dev, ops.LANCZOS)
    stack, _ = ops.resnet50_features(ctx, rn_in, is_bgr=True)
    vit = ops.vitb16_features(ctx, vt_in, is_bgr=True)
    assert max(seg_err(stack.cpu().numpy(), g["full_resnet"], SEGMENTS)) <= 1e-2
    assert max(seg_err(vit.cpu().numpy(), g["full_vit"], [768, 768, 768])) <= 1e-2


def _many_images(n, seed=5):
    """n distinct 224x224 BGR images: smooth synthetic frames, their rolls / flips, and white noise."""
    fr, nx = synth.make_clip(seed, 224, 224, 12)
    base = np.concatenate([fr, nx])                                        # 24 images
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        b = base[i % len(base)]
        k = i // len(base)
        if k % 4 == 1:
            b = np.roll(b, (7 * k, 13 * k), axis=(0, 1))
        elif k % 4 == 2:
            b = b[::-1, :, ::-1]
        elif k % 4 == 3:
            b = rng.integers(0, 256, b.shape, dtype=np.uint8)
        out.append(np.ascontiguousarray(b))
    return np.stack(out)


def test_backbones_at_the_bench_shape(ctx):
    """parity at the batch the 1080p benchmark runs (4 clips x 66 = 264 images per pass: 204 row-tile
    pairs, 26 raster groups, 9/25/34 waves) - tcgen05 vs the SIMT check path on all rows, 8 sampled rows vs the fp32
    oracle, and batch invariance against single-image calls."""
    from relax_vqa_b200 import ops
    rsd, vsd = weights.seeded_resnet50_state_dict(1234), weights.seeded_vitb16_state_dict(4321)
    images = _many_images(264)
    dev = torch.from_numpy(images).cuda()
    ctx.set_gemm_impl(1)
    chk_stack, chk_pool = ops.resnet50_features(ctx, dev, want_stack=True, want_pool=True)
    chk_vit = ops.vitb16_features(ctx, dev)
    ctx.set_gemm_impl(0)
    stack, pool = ops.resnet50_features(ctx, dev, want_stack=True, want_pool=True)
    vit = ops.vitb16_features(ctx, dev)
    torch.cuda.synchronize()
    e_rn = seg_err(stack.cpu().numpy(), chk_stack.cpu().numpy(), SEGMENTS)
    e_vt = seg_err(vit.cpu().numpy(), chk_vit.cpu().numpy(), [768] * 3)
    print("264 images, tcgen05 vs SIMT check: resnet", max(e_rn), "vit", max(e_vt))
    assert max(e_rn) < 2e-3 and max(e_vt) < 3e-3
    assert max(seg_err(pool.cpu().numpy(), chk_pool.cpu().numpy(), [2048, 1, 1, 1])) < 2e-3
    rows = [0, 1, 37, 100, 131, 200, 262, 263]
    x = OB.resnet_preprocess(images[rows][..., ::-1])
    assert max(seg_err(stack[rows].cpu().numpy(), OB.resnet50_layerstack(rsd, x), SEGMENTS)) <= 1e-2
    assert max(seg_err(pool[rows].cpu().numpy(), OB.resnet50_pool(rsd, x), [2048, 1, 1, 1])) <= 1e-2
    assert max(seg_err(vit[rows].cpu().numpy(), OB.vit_pool(vsd, OB.vit_preprocess(images[rows][..., ::-1])), [768] * 3)) <= 1e-2
    for r in (0, 131, 263):                                                # batch invariance against B = 1, bit for bit
        one_s, one_p = ops.resnet50_features(ctx, dev[r:r + 1].contiguous(), want_stack=True, want_pool=True)
        assert torch.equal(one_s[0], stack[r]) and torch.equal(one_p[0], pool[r])
        assert torch.equal(ops.vitb16_features(ctx, dev[r:r + 1].contiguous())[0], vit[r])


def test_backbones_split_passes_above_512_images(ctx):
    """B = 520 executes the > 512-image split (two balanced passes of 260): every row equals the row of a smaller batch,
    bit for bit, and sampled rows match the fp32 oracle."""
    from relax_vqa_b200 import ops
    rsd, vsd = weights.seeded_resnet50_state_dict(1234), weights.seeded_vitb16_state_dict(4321)
    images = _many_images(520, seed=6)
    dev = torch.from_numpy(images).cuda()
    stack, pool = ops.resnet50_features(ctx, dev, want_stack=True, want_pool=True)
    vit = ops.vitb16_features(ctx, dev)
    for lo, hi in ((0, 7), (255, 265), (513, 520)):                        # windows that straddle the pass boundary (260) and the ends
        s2, p2 = ops.resnet50_features(ctx, dev[lo:hi].contiguous(), want_stack=True, want_pool=True)
        assert torch.equal(s2, stack[lo:hi]) and torch.equal(p2, pool[lo:hi])
        assert torch.equal(ops.vitb16_features(ctx, dev[lo:hi].contiguous()), vit[lo:hi])
    rows = [0, 259, 260, 519]
    x = OB.resnet_preprocess(images[rows][..., ::-1])
    assert max(seg_err(stack[rows].cpu().numpy(), OB.resnet50_layerstack(rsd, x), SEGMENTS)) <= 1e-2
    assert max(seg_err(vit[rows].cpu().numpy(), OB.vit_pool(vsd, OB.vit_preprocess(images[rows][..., ::-1])), [768] * 3)) <= 1e-2
