"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/gen_golden.py

The reference is imported in place through the five shims of SURVEY.md 8(c):
  1. sys.modules stubs for matplotlib / seaborn / ipywidgets (imported, never used on the path);
  2. torchvision.models.resnet50 / vgg16 factories returning seeded models (both module-global
     ResNet copies share weights);
  3. torch.hub.load_state_dict_from_url returning the seeded ViT-B/16 state dict;
  4. cwd = scratch dir with a writable utils/ (the reference logs to ./utils/log_debug.log and
     writes ../visualisation, ../features);
  5. demo_test.load patched to add `_fill_dtype` to the sklearn-1.3.2 imputer pickle.
Outputs (small .npz files) are committed; tests compare the oracle and the CUDA path to them.
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RELAXVQA_REF", "/root/reference")
sys.path.insert(0, ROOT)

from relax_vqa_b200 import synth, weights  # noqa: E402

RESNET_SEED, VIT_SEED, HEAD_SEED = 1234, 4321, 99


def install_shims():
    import torch
    import torchvision
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "seaborn", "ipywidgets"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    real_resnet50 = torchvision.models.resnet50

    def seeded_resnet50(*a, **k):
        m = real_resnet50(weights=None)
        m.load_state_dict(weights.seeded_resnet50_state_dict(RESNET_SEED))
        return m

    def tiny_vgg16(*a, **k):          # imported at module import by the drivers, never run
        return torch.nn.Sequential()

    torchvision.models.resnet50 = seeded_resnet50
    torchvision.models.vgg16 = tiny_vgg16
    torch.hub.load_state_dict_from_url = lambda *a, **k: weights.seeded_vitb16_state_dict(VIT_SEED)


def main():
    import cv2
    import joblib
    import torch
    scratch = tempfile.mkdtemp(prefix="relaxvqa_golden_")
    work = os.path.join(scratch, "src")
    os.makedirs(os.path.join(work, "utils"))
    os.chdir(work)
    sys.path.insert(0, os.path.join(REF, "src"))
    install_shims()
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import demo_test
    import main_fragment_layerstack as mfl

    real_load = demo_test.load

    def patched_load(path):
        obj = real_load(path)
        if hasattr(obj, "statistics_") and not hasattr(obj, "_fill_dtype"):
            obj._fill_dtype = obj._fit_dtype
        return obj

    demo_test.load = patched_load

    # ---- tiny synthetic clip written with the reference's PNG naming ------------------
    H, W, T = 272, 480, 3
    vid = "synth0"
    frames, nexts = synth.make_clip(7, H, W, T)
    base = os.path.join(scratch, "video_sampled_frame", "original_sampled_frame")
    d_frames = os.path.join(base, "test_sampled_frames")
    d_frag = os.path.join(base, "test_sampled_fragment")
    os.makedirs(d_frames)
    os.makedirs(d_frag)
    for t in range(T):
        cv2.imwrite(os.path.join(d_frames, f"{vid}_{t + 1}.png"), frames[t])
        cv2.imwrite(os.path.join(d_frag, f"{vid}_{t + 1}.png"), frames[t])
        cv2.imwrite(os.path.join(d_frag, f"{vid}_{t + 1}_next.png"), nexts[t])

    # ---- (1) fragment stages through the reference's own helpers ----------------------
    frag = {}
    for t in range(T):
        a, b = frames[t], nexts[t]
        residual = cv2.absdiff(b, a)
        sums = mfl.get_patch_diff(residual, 16)
        _, diff_frag, positions = mfl.process_patches("x.png", "frame_diff", residual, 16, 224, 196)
        ori = mfl.get_original_frame_patches(a, positions, 16, 224)
        flow = cv2.calcOpticalFlowFarneback(cv2.cvtColor(a, cv2.COLOR_BGR2GRAY), cv2.cvtColor(b, cv2.COLOR_BGR2GRAY),
                                            None, 0.5, 3, 15, 3, 5, 1.2, 0)
        rgb = mfl.flow_to_rgb(flow)
        _, flow_frag, fpos = mfl.process_patches("x.png", "optical_flow", rgb, 16, 224, 196)
        merged = mfl.merge_fragments(diff_frag, flow_frag)
        frag.update({f"sums{t}": sums, f"pos{t}": np.array(positions, dtype=np.int32), f"diff_frag{t}": diff_frag,
                     f"ori_frag{t}": ori, f"flow{t}": flow.astype(np.float16), f"flow_rgb{t}": rgb,
                     f"flow_pos{t}": np.array(fpos, dtype=np.int32), f"flow_frag{t}": flow_frag,
                     f"merged{t}": merged})
    np.savez_compressed(os.path.join(HERE, "ref_fragments_synth.npz"), H=H, W=W, T=T, clip_seed=7, **frag)

    # ---- (2) per-image backbone features + (3) full video vector, reference call stack --
    from main_layer_stack import get_deep_feature as gdf_ls, process_video_feature as pvf_ls
    from main_fragment_pool import get_deep_feature as gdf_fp, process_video_feature as pvf_fp
    qp = "original"
    rn, vt = [], []
    for t in range(T):
        p = os.path.join(d_frames, f"{vid}_{t + 1}.png")
        rn.append(gdf_ls("resnet50", vid, p, qp)[2])
        vt.append(gdf_ls("vit", vid, p, qp)[2])
    full_resnet = pvf_ls(rn, "resnet50")
    full_vit = pvf_ls(vt, "vit")
    o_rn, r_rn, o_vt, r_vt = [], [], [], []
    for t in range(T):
        op = os.path.join(d_frag, f"{vid}_{t + 1}_ori_frag.png")
        mp = os.path.join(d_frag, f"{vid}_{t + 1}_residual_merged_frag.png")
        cv2.imwrite(op, frag[f"ori_frag{t}"])
        cv2.imwrite(mp, frag[f"merged{t}"])
        o_rn.append(mfl.get_deep_feature("resnet50", vid, op, qp, "layer_stack")[2])
        r_rn.append(mfl.get_deep_feature("resnet50", vid, mp, qp, "pool")[2])
        o_vt.append(gdf_fp("vit", vid, op, qp, "pool")[2])
        r_vt.append(gdf_fp("vit", vid, mp, qp, "pool")[2])
        os.remove(op)
        os.remove(mp)
    frag_resnet = mfl.concatenate_features(mfl.process_video_feature(o_rn, "resnet50", "layer_stack"),
                                           mfl.process_video_feature(r_rn, "resnet50", "pool"))
    frag_vit = mfl.concatenate_features(pvf_fp(o_vt, "vit"), pvf_fp(r_vt, "vit"))
    vec = np.concatenate([np.mean(full_resnet, 0), np.mean(full_vit, 0), np.mean(frag_resnet, 0), np.mean(frag_vit, 0)])
    assert vec.shape == (35203,), vec.shape

    # ---- (4) the unmodified evaluate_video_quality end to end -------------------------
    save_path = os.path.join(scratch, "model")
    os.makedirs(os.path.join(save_path, "scaler"))
    for f in ("konvid_1k_imputer.pkl", "konvid_1k_scaler.pkl"):
        shutil.copy(os.path.join(REF, "model", "scaler", f), os.path.join(save_path, "scaler", f))
    torch.save(weights.seeded_head_state_dict(HEAD_SEED, swa_format=True),
               os.path.join(save_path, "lsvq_train_relaxvqa_byrmse_trained_median_model_param_onLSVQ_TEST.pth"))
    cfg = dict(device=torch.device("cpu"), model_name="Mlp", layer_name="pool", select_criteria="byrmse",
               train_data_name="lsvq_train", is_finetune=False, save_path=save_path, video_type="konvid_1k",
               video_name=vid, qp=qp, video_width=W, video_height=H, pixfmt="yuv420p", framerate=29.97)
    score = demo_test.evaluate_video_quality(cfg)
    imp = patched_load(os.path.join(save_path, "scaler", "konvid_1k_imputer.pkl"))
    np.savez_compressed(os.path.join(HERE, "ref_video_synth.npz"), H=H, W=W, T=T, clip_seed=7,
                        resnet_seed=RESNET_SEED, vit_seed=VIT_SEED, head_seed=HEAD_SEED,
                        full_resnet=full_resnet.astype(np.float32), full_vit=full_vit.astype(np.float32),
                        frag_resnet=frag_resnet.astype(np.float32), frag_vit=frag_vit.astype(np.float32),
                        vector=vec.astype(np.float32), score=np.float64(score))
    # scaler / imputer attributes (the pickles themselves stay in the reference tree)
    sc = joblib.load(os.path.join(REF, "model", "scaler", "konvid_1k_scaler.pkl"))
    np.savez_compressed(os.path.join(HERE, "konvid_1k_scaler_imputer.npz"),
                        imputer_mean=imp.statistics_.astype(np.float64), scale=sc.scale_.astype(np.float64),
                        minv=sc.min_.astype(np.float64))
    print("score", score, "vector", vec[:4], vec.shape)
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
