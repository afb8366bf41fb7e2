"""Oracle (CPU) for the whole hot path: sampled frame pairs -> 35,203 features -> score.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Restates src/demo_test.py:51-219 on in-memory frames (the reference hands PNG files between
stages; PNG is lossless, so arrays are equivalent).  `frames`/`nexts` are BGR uint8 as
cv2.imread returns them; PIL opens the same PNGs as RGB, hence the [..., ::-1] flips.
"""
import numpy as np

from . import backbones as B
from . import farneback as FB
from . import fragments as FR
from . import head as HD
from . import resize as RS


def pair_fragments(frame, nxt, flow_fn=FB.farneback):
    """Per-pair image stages (src/demo_test.py:104-135) -> dict of uint8 arrays (BGR)."""
    residual = FR.absdiff(nxt, frame)
    diff_frag, pos, sums = FR.process_patches(residual)
    ori_frag = FR.gather_fragment(frame, pos)
    flow = flow_fn(FR.bgr2gray(frame), FR.bgr2gray(nxt))
    flow_rgb = FB.flow_to_rgb(flow)
    flow_frag, flow_pos, flow_sums = FR.process_patches(flow_rgb)
    merged = FR.merge_fragments(diff_frag, flow_frag)
    return dict(residual=residual, sums=sums, positions=pos, diff_frag=diff_frag, ori_frag=ori_frag,
                flow=flow, flow_rgb=flow_rgb, flow_sums=flow_sums, flow_positions=flow_pos,
                flow_frag=flow_frag, merged_frag=merged)


def video_feature_blocks(frames, nexts, resnet_sd, vit_sd, flow_fn=FB.farneback, batch=8):
    """-> dict of per-frame matrices: full_resnet (Tf,13120), full_vit (Tf,2304),
    frag_resnet (Tp,15171), frag_vit (Tp,4608)  (src/demo_test.py:80-161)."""
    def batched(fn, imgs):
        return np.concatenate([fn(np.stack(imgs[i:i + batch])) for i in range(0, len(imgs), batch)], axis=0)

    full_rn_in = [RS.resize(f[..., ::-1], 224, 224, RS.BILINEAR) for f in frames]
    full_vit_in = [RS.resize(f[..., ::-1], 224, 224, RS.LANCZOS) for f in frames]
    full_resnet = batched(lambda u: B.resnet50_layerstack(resnet_sd, B.resnet_preprocess(u)), full_rn_in)
    full_vit = batched(lambda u: B.vit_pool(vit_sd, B.vit_preprocess(u)), full_vit_in)
    pairs = [pair_fragments(f, n, flow_fn) for f, n in zip(frames, nexts)]
    ori = [p["ori_frag"][..., ::-1] for p in pairs]
    mer = [p["merged_frag"][..., ::-1] for p in pairs]
    frag_resnet = np.concatenate([
        batched(lambda u: B.resnet50_layerstack(resnet_sd, B.resnet_preprocess(u)), ori),
        batched(lambda u: B.resnet50_pool(resnet_sd, B.resnet_preprocess(u)), mer)], axis=1)
    frag_vit = np.concatenate([
        batched(lambda u: B.vit_pool(vit_sd, B.vit_preprocess(u)), ori),
        batched(lambda u: B.vit_pool(vit_sd, B.vit_preprocess(u)), mer)], axis=1)
    return dict(full_resnet=full_resnet, full_vit=full_vit, frag_resnet=frag_resnet, frag_vit=frag_vit,
                pairs=pairs)


def video_vector(blocks):
    """Temporal mean of each block and concatenation -> (35203,) fp32 (src/demo_test.py:171-175)."""
    return np.concatenate([np.mean(blocks[k], axis=0) for k in
                           ("full_resnet", "full_vit", "frag_resnet", "frag_vit")])


def predict(vec, head_sd, imputer_mean, scale, minv, video_type="konvid_1k", is_finetune=False):
    x = HD.impute_scale(vec.reshape(1, -1), imputer_mean, scale, minv)
    pred = float(HD.mlp_forward(head_sd, x.astype(np.float32))[0])
    return HD.rescale_score(pred, video_type, is_finetune)
