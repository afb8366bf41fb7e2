"""GPU parity of the frame sampler: raw yuv420p -> BGR on the device vs the swscale-pinned oracle
(bit-exact), container decode vs sequential cv2 decode, the reference-named PNG writers, and evaluate_video_quality fed
by a video file instead of pre-sampled PNGs."""
import glob
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import sampler as S
from relax_vqa_b200 import synth, weights

pytestmark = pytest.mark.gpu


def _write_yuv(path, H, W, n, seed=3):
    """n frames of a drifting synthetic texture as planar yuv420p (converted with cv2: any legal planes will do)."""
    if H >= 64:
        fr, nx = synth.make_clip(seed, H, W, 2)
    else:                                                   # too small for the synthetic clip generator: noise
        rng = np.random.default_rng(seed)
        fr, nx = rng.integers(0, 256, (2, 2, H, W, 3), dtype=np.uint8)
    frames = []
    with open(path, "wb") as f:
        for i in range(n):
            img = np.roll(fr[i % 2] if i % 3 else nx[i % 2], (2 * i, 3 * i), axis=(0, 1))
            yuv = cv2.cvtColor(img, cv2.COLOR_BGR2YUV_I420)            # (H*3/2, W) planar
            f.write(yuv.tobytes())
            frames.append(yuv.reshape(-1).copy())
    return np.stack(frames)


@pytest.mark.parametrize("hw", [(272, 480), (1080, 1920), (38, 100)])
def test_yuv420p_sampler_bit_exact(tmp_path, hw):
    from relax_vqa_b200 import ops, video_frames_extract as vfe
    H, W = hw
    n, k = (31, 14) if H < 1000 else (16, 14)
    path = str(tmp_path / "clip.yuv")
    raw = _write_yuv(path, H, W, n)
    clip = vfe.sample_yuv420p(path, W, H, k)
    fr, nx = S.sample_yuv420p(path, W, H, k)
    assert clip.frames.shape == fr.shape and clip.nexts.shape == nx.shape
    assert np.array_equal(clip.frames.cpu().numpy(), fr) and np.array_equal(clip.nexts.cpu().numpy(), nx)
    # every byte value through the converter (random planes, out-of-range included)
    rng = np.random.default_rng(0)
    planes = rng.integers(0, 256, (2, H * W * 3 // 2), dtype=np.uint8)
    got = ops.yuv420p_to_bgr(torch.from_numpy(planes).cuda(), H, W).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], S.yuv420p_to_bgr(*S.split_planes(planes[i], H, W)))
    with pytest.raises(ValueError):
        vfe.sample_yuv420p(path, W, H, k, pixfmt="yuv422p")


def test_container_sampler_and_png_writers(tmp_path):
    from relax_vqa_b200 import video_frames_extract as vfe
    from relax_vqa_b200.extractor import vf_extract
    H, W, n, k = 144, 256, 33, 14
    fr, nx = synth.make_clip(8, H, W, 2)
    path = str(tmp_path / "vid01.avi")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 29.97, (W, H))
    assert w.isOpened()
    for i in range(n):
        w.write(np.roll(fr[i % 2], (i, 2 * i), axis=(0, 1)))
    w.release()
    cap = cv2.VideoCapture(path)
    dec = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        dec.append(f)
    assert len(dec) == n
    clip = vfe.sample_video(path, k)
    full, nxt = S.selected_indices(n, k)
    assert np.array_equal(clip.frames.cpu().numpy(), np.stack([dec[i] for i in full]))
    assert np.array_equal(clip.nexts.cpu().numpy(), np.stack([dec[i] for i in nxt[:len(full)]]))
    out = str(tmp_path / "sampled")
    vf_extract.process_video_residual("konvid_1k", "vid01", k, path, out, W, H, "yuv420p", 29.97)       # ref names, :9, :61
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(out, "*.png")))
    assert names == sorted([f"vid01_{i + 1}.png" for i in range(3)] + [f"vid01_{i + 1}_next.png" for i in range(3)])
    assert np.array_equal(cv2.imread(os.path.join(out, "vid01_2_next.png")), dec[15])
    out2 = str(tmp_path / "sampled_full")
    vfe.process_video("konvid_1k", "vid01", k, path, out2, W, H, "yuv420p", 29.97)
    assert sorted(os.listdir(out2)) == [f"vid01_{i + 1}.png" for i in range(3)]


def test_evaluate_video_quality_from_a_raw_video(tmp_path, golden_dir):
    """demo_test.evaluate_video_quality with config['video_path'] (live_qualcomm-style raw yuv) == the same entry point on
    the PNGs process_video / process_video_residual write for that video (PNG is lossless): sampler -> engine hand-off."""
    import joblib
    from sklearn.impute import SimpleImputer
    from sklearn.preprocessing import MinMaxScaler
    from relax_vqa_b200 import demo_test, runtime, video_frames_extract as vfe
    H, W, n, fps = 272, 480, 31, 29.97
    path = str(tmp_path / "q01.yuv")
    _write_yuv(path, H, W, n, seed=5)
    s = np.load(os.path.join(golden_dir, "konvid_1k_scaler_imputer.npz"))
    save_path = tmp_path / "model"
    (save_path / "scaler").mkdir(parents=True)
    imp = SimpleImputer(strategy="mean"); imp.statistics_ = s["imputer_mean"]
    sc = MinMaxScaler(); sc.scale_, sc.min_ = s["scale"], s["minv"]
    joblib.dump(imp, save_path / "scaler" / "live_qualcomm_imputer.pkl")
    joblib.dump(sc, save_path / "scaler" / "live_qualcomm_scaler.pkl")
    torch.save(weights.seeded_head_state_dict(99, swa_format=True),
               save_path / "lsvq_train_relaxvqa_byrmse_trained_median_model_param_onLSVQ_TEST.pth")
    runtime.configure(0, weights.seeded_resnet50_state_dict(1234), weights.seeded_vitb16_state_dict(4321))
    cfg = dict(device=torch.device("cuda"), model_name="Mlp", layer_name="pool", select_criteria="byrmse", train_data_name="lsvq_train",
               is_finetune=False, save_path=str(save_path), video_type="live_qualcomm", video_name="q01", qp="original",
               video_width=W, video_height=H, pixfmt="yuv420p", framerate=fps)
    direct = demo_test.evaluate_video_quality(dict(cfg, video_path=path))
    root = tmp_path / "sampled"
    vfe.process_video("live_qualcomm", "q01", int(fps / 2), path, str(root / "test_sampled_frames"), W, H, "yuv420p", fps)
    vfe.process_video_residual("live_qualcomm", "q01", int(fps / 2), path, str(root / "test_sampled_fragment"), W, H, "yuv420p", fps)
    via_png = demo_test.evaluate_video_quality(dict(cfg, sampled_root=str(root)))
    print("score from the raw video", direct, "from its sampled PNGs", via_png)
    assert direct == via_png and np.isfinite(direct)
    runtime.configure(0, None, None)
