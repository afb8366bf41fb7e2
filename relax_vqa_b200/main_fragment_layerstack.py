"""Drop-in for the hot-path functions of the reference's src/main_fragment_layerstack.py (same names,
argument meaning and return types; numpy in / numpy out), executed by libb200vqa on the GPU."""
import os

import cv2
import numpy as np

from . import ops, runtime
from .runtime import PooledFrame

RESNET_LAYERS = ['resnet50.conv1',
                 'resnet50.layer1[0]', 'resnet50.layer1[1]', 'resnet50.layer1[2]',
                 'resnet50.layer2[0]', 'resnet50.layer2[1]', 'resnet50.layer2[2]', 'resnet50.layer2[3]',
                 'resnet50.layer3[0]', 'resnet50.layer3[1]', 'resnet50.layer3[2]', 'resnet50.layer3[3]',
                 'resnet50.layer4[0]', 'resnet50.layer4[1]', 'resnet50.layer4[2]']


def _require_patch16(patch_size, target_size=224, top_n=196):
    if patch_size != 16 or target_size != 224 or top_n != 196:
        raise ValueError("libb200vqa implements the reference's constants only: patch 16, target 224, top_n 196 "
                         "(src/main_fragment_layerstack.py:297-299)")


def get_patch_diff(residual_frame, patch_size):
    """ref :177-189 -> float64 (gh, gw) of exact patch sums."""
    _require_patch16(patch_size)
    sums = ops.patchsum(runtime.to_dev(residual_frame[None]))
    return sums[0].cpu().numpy().astype(np.float64)


def extract_important_patches(residual_frame, diff, patch_size=16, target_size=224, top_n=196):
    """ref :191-210 -> (fragment (224,224,3) uint8, positions list[(y, x)] in raster order)."""
    _require_patch16(patch_size, target_size, top_n)
    img = runtime.to_dev(residual_frame[None])
    d = np.asarray(diff, dtype=np.float64)
    if d.size and np.all(d >= 0) and np.all(d == np.floor(d)) and d.max() < 2 ** 31:
        keys = d.astype(np.int32)                      # what get_patch_diff returns: exact integer sums
    else:
        # any other float map (the reference accepts one): dense ranks keep the order and the ties of the values,
        # so the kernel's (value desc, index asc) rule picks what np.argsort(-diff, kind="stable") picks
        keys = np.unique(d, return_inverse=True)[1].reshape(d.shape).astype(np.int32)
    sums = runtime.to_dev(keys[None])
    pos, cnt = ops.topk_patches(sums)
    frag, _ = ops.gather_fragments(img, None, pos, cnt, want_ori=True, want_diff=False)
    k = int(cnt[0])
    return frag[0].cpu().numpy(), [tuple(p) for p in pos[0, :k].cpu().numpy().tolist()]


def get_original_frame_patches(original_frame, positions, patch_size, target_size):
    """ref :212-230."""
    _require_patch16(patch_size, target_size)
    import torch
    pos = torch.full((1, 196, 2), -1, dtype=torch.int32)
    if len(positions):
        pos[0, :len(positions)] = torch.tensor(positions, dtype=torch.int32)
    cnt = torch.tensor([len(positions)], dtype=torch.int32)
    dev = runtime.engine().device
    frag, _ = ops.gather_fragments(runtime.to_dev(original_frame[None]), None, pos.to(dev), cnt.to(dev), want_diff=False)
    return frag[0].cpu().numpy()


def process_patches(original_path, residual_name, residual, patch_size, target_size, top_n):
    """ref :232-240."""
    diff = get_patch_diff(residual, patch_size)
    imp_patches, positions = extract_important_patches(residual, diff, patch_size, target_size, top_n)
    suffix = '_residual_imp.png' if residual_name == 'frame_diff' else '_residual_of_imp.png'
    return original_path.replace('.png', suffix), imp_patches, positions


def flow_to_rgb(flow):
    """ref :162-175 -> (H, W, 3) uint8 (BGR, like the reference)."""
    rgb, _, _ = ops.flow_to_rgb(runtime.to_dev(np.asarray(flow, dtype=np.float32)[None]), want_rgb=True, want_sums=False)
    return rgb[0].cpu().numpy()


def calc_optical_flow_farneback(img_original, img_next):
    """cv2.calcOpticalFlowFarneback(gray(img_original), gray(img_next), None, 0.5, 3, 15, 3, 5, 1.2, 0) (ref :313-315)."""
    r = ops.absdiff_patchsum(runtime.to_dev(img_original[None]), runtime.to_dev(img_next[None]))
    return ops.farneback(runtime.engine().ctx, r["gray0"], r["gray1"])[0].cpu().numpy()


def merge_fragments(diff_fragment, flow_fragment):
    """ref :242-245."""
    return ops.merge_fragments(runtime.to_dev(diff_fragment), runtime.to_dev(flow_fragment)).cpu().numpy()


def concatenate_features(original_feature, residual_feature):
    """ref :247-248."""
    return np.concatenate((original_feature, residual_feature), axis=-1)


def _load_bgr(image_path):
    img = cv2.imread(image_path, cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(image_path)
    return img


def _resnet_input(img_bgr):
    """transforms.Resize((224,224)) on the PIL image (ref visualise_resnet.py:40-47)."""
    eng = runtime.engine()
    return ops.resize_pil(eng.ctx, runtime.to_dev(img_bgr[None]), ops.BILINEAR)


def _vit_input(img_bgr):
    """img.resize((224,224), LANCZOS) when the size differs (ref visualise_vit_layer.py:466-470)."""
    eng = runtime.engine()
    return ops.resize_pil(eng.ctx, runtime.to_dev(img_bgr[None]), ops.LANCZOS)


def get_deep_feature(network_name, video_name, image_path, qp, layer_name):
    """ref :83-121.  Returns (png_path, npy_path, frame_feature).

    Default (fast path): frame_feature is the already-pooled vector for this image (PooledFrame): (13120,) for
    resnet50/layer_stack, (2048,) for resnet50/pool, (2304,) for vit.
    With runtime.configure(return_maps=True) it has the reference's own types: dict[layer name -> (C,H,W) float32] for
    resnet50/layer_stack, (2048,1,1) for resnet50/pool, (196,768) for vit.
    process_video_feature below consumes either and yields the reference's (T, D) rows."""
    png_path = f'../visualisation/{network_name}/{video_name}/'
    npy_path = f'../features/{network_name}/{video_name}/'
    eng = runtime.engine()
    img = _load_bgr(image_path)
    if runtime.return_maps():
        if network_name == 'resnet50':
            maps = ops.resnet50_maps(eng.ctx, _resnet_input(img), is_bgr=True)
            if layer_name == 'layer_stack':
                return png_path, npy_path, {name: m[0].cpu().numpy() for name, m in zip(RESNET_LAYERS, maps)}
            if layer_name == 'pool':
                # the avgpool hook (visualise_resnet_layer.py:62-102): spatial mean of layer4[2]'s output
                return png_path, npy_path, maps[-1][0].mean(dim=(1, 2)).reshape(2048, 1, 1).cpu().numpy()
        elif network_name == 'vit':
            return png_path, npy_path, ops.vitb16_tokens(eng.ctx, _vit_input(img), is_bgr=True)[0].cpu().numpy()
        raise ValueError(f"unsupported network/layer on the B200 hot path: {network_name}/{layer_name}")
    if network_name == 'resnet50':
        stack, pool = ops.resnet50_features(eng.ctx, _resnet_input(img), is_bgr=True, want_stack=True, want_pool=True)
        if layer_name == 'layer_stack':
            return png_path, npy_path, PooledFrame(stack[0].cpu().numpy(), 'resnet50_layer_stack')
        if layer_name == 'pool':
            return png_path, npy_path, PooledFrame(pool[0, :2048].cpu().numpy(), 'resnet50_avgpool')
    elif network_name == 'vit':
        vit = ops.vitb16_features(eng.ctx, _vit_input(img), is_bgr=True)
        return png_path, npy_path, PooledFrame(vit[0].cpu().numpy(), 'vit_pool')
    raise ValueError(f"unsupported network/layer on the B200 hot path: {network_name}/{layer_name}")


def process_video_feature(video_feature, network_name, layer_name):
    """ref :124-160 -> (T, 13120) for 'layer_stack', (T, 2051) for 'pool'.  Each element of `video_feature` is what
    get_deep_feature returned: a PooledFrame (fast path) or the reference's raw type (return_maps=True), pooled here
    the way the reference pools it (per-layer np.mean over (H, W), :131-140; hstack of [v, mean, max, std], :141-149)."""
    rows = []
    for frame in video_feature:
        if layer_name == 'layer_stack':
            if isinstance(frame, dict):
                rows.append(np.concatenate([np.mean(np.asarray(a, dtype=np.float32), axis=(1, 2)) for a in frame.values()]))
            else:
                rows.append(np.asarray(frame, dtype=np.float32))
        else:
            v = np.squeeze(np.asarray(frame, dtype=np.float32))
            rows.append(np.hstack([v, np.mean(v, axis=0), np.max(v, axis=0), np.std(v, axis=0)]).astype(np.float32))
    return np.array(rows)
