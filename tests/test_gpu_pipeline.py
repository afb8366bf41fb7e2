"""End-to-end GPU parity: sampled pairs -> 35,203 vector -> MOS through the public Engine API and
through the reference-named drop-in functions, against (a) the CPU oracle on the same inputs and
(b) the UNMODIFIED reference's outputs (tests/golden/ref_video_synth.npz).

Tolerances (north_star): stacked features within 1e-2 (per segment, max|a-b|/rms), MOS within 0.01,
frame-diff fragments bit-exact; the flow-dependent merged fragment is a float-tolerance stage: report
its pixel mismatch rate and bound it."""
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import fragments as F
from oracle import pipeline as P
from relax_vqa_b200 import synth, weights

pytestmark = pytest.mark.gpu

BLOCK_SEGS = {
    "full_resnet": [64, 256, 256, 256, 512, 512, 512, 512, 1024, 1024, 1024, 1024, 2048, 2048, 2048],
    "full_vit": [768, 768, 768],
    "frag_resnet": [64, 256, 256, 256, 512, 512, 512, 512, 1024, 1024, 1024, 1024, 2048, 2048, 2048, 2048, 3],
    "frag_vit": [768] * 6,
}
VEC_SEGS = BLOCK_SEGS["full_resnet"] + BLOCK_SEGS["full_vit"] + BLOCK_SEGS["frag_resnet"] + BLOCK_SEGS["frag_vit"]


def _record(name, row):
    """Observed tolerances are appended to gpurun_out/parity_observed.jsonl (copied into profiles/ per round)."""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_observed.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **row)) + "\n")
    except OSError:
        pass


def seg_err(a, b, widths):
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    out, o = [], 0
    for w in widths:
        ref = b[:, o:o + w]
        out.append(np.abs(a[:, o:o + w] - ref).max() / max(np.sqrt(np.mean(ref ** 2)), 1e-12))
        o += w
    assert o == b.shape[1]
    return out


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_video_synth.npz"))


@pytest.fixture(scope="module")
def engine(golden, golden_dir):
    from relax_vqa_b200.engine import Engine
    s = np.load(os.path.join(golden_dir, "konvid_1k_scaler_imputer.npz"))
    e = Engine(0, weights.seeded_resnet50_state_dict(int(golden["resnet_seed"])), weights.seeded_vitb16_state_dict(int(golden["vit_seed"])),
               weights.seeded_head_state_dict(int(golden["head_seed"]), swa_format=True), s["imputer_mean"], s["scale"], s["minv"])
    yield e
    e.close()


def test_video_vector_and_score_vs_reference_golden(engine, golden):
    from relax_vqa_b200.engine import Clip
    g = golden
    fr, nx = synth.make_clip(int(g["clip_seed"]), int(g["H"]), int(g["W"]), int(g["T"]))
    clip = Clip(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda())
    b = engine.extract_blocks([clip])
    got = dict(full_resnet=b["full_resnet"], full_vit=b["full_vit"],
               frag_resnet=torch.cat([b["frag_stack"], b["frag_pool"]], dim=1),
               frag_vit=torch.cat([b["frag_vit_ori"], b["frag_vit_mer"]], dim=1))
    for k, segs in BLOCK_SEGS.items():
        e = seg_err(got[k].cpu().numpy(), g[k], segs)
        print(k, "max seg err", max(e))
        assert max(e) <= 1e-2, (k, e)
    feats, score = engine.predict([clip], "konvid_1k")
    assert feats.shape == (1, 35203)
    assert max(seg_err(feats.cpu().numpy(), g["vector"][None], VEC_SEGS)) <= 1e-2
    print("score", float(score[0]), "reference", float(g["score"]))
    assert abs(float(score[0]) - float(g["score"])) <= 0.01


def test_fragments_vs_oracle_and_reference(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_fragments_synth.npz"))
    fr, nx = synth.make_clip(int(g["clip_seed"]), int(g["H"]), int(g["W"]), int(g["T"]))
    out = engine.fragments(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda(), keep_intermediates=True)
    for t in range(int(g["T"])):
        assert np.array_equal(out["sums"][t].cpu().numpy().astype(np.float64), g[f"sums{t}"])
        assert np.array_equal(out["positions"][t].cpu().numpy(), g[f"pos{t}"])
        assert np.array_equal(out["diff_frag"][t].cpu().numpy(), g[f"diff_frag{t}"])          # bit-exact
        assert np.array_equal(out["ori_frag"][t].cpu().numpy(), g[f"ori_frag{t}"])            # bit-exact
        flow_err = np.abs(out["flow"][t].cpu().numpy() - g[f"flow{t}"].astype(np.float32))
        assert flow_err.max() < 2e-2                                                           # golden flow is stored as fp16
        mism = (out["merged_frag"][t].cpu().numpy() != g[f"merged{t}"]).any(-1).mean()
        pos_sym = len(set(map(tuple, out["flow_positions"][t].cpu().numpy().tolist())) ^ set(map(tuple, g[f"flow_pos{t}"].tolist())))
        print(f"pair {t}: merged-fragment pixel mismatch rate {mism:.5f}, flow patch-set symmetric difference {pos_sym}")
        _record("merged_fragment_vs_reference", dict(pair=t, pixel_mismatch_rate=float(mism), flow_patch_set_symmetric_difference=int(pos_sym),
                                                      flow_max_abs_err_px_vs_fp16_golden=float(flow_err.max())))
        assert mism < 2e-3 and pos_sym <= 2          # observed r2 (cv2-exact colouring): <= 4e-5 and 0, profiles/r2_parity_observed.jsonl


def test_multi_clip_batch_equals_single(engine):
    from relax_vqa_b200.engine import Clip
    clips = []
    for seed, hw in ((1, (272, 480)), (2, (144, 256)), (3, (272, 480))):
        fr, nx = synth.make_clip(seed, hw[0], hw[1], 2)
        clips.append(Clip(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda()))
    all_feats = engine.extract(clips)
    for i, c in enumerate(clips):
        assert torch.equal(engine.extract([c])[0], all_feats[i])        # batch-invariant, bit for bit


def test_dropin_functions_match_oracle(engine, tmp_path):
    from relax_vqa_b200 import main_fragment_layerstack as mfl
    from relax_vqa_b200 import runtime
    runtime._engine = engine
    fr, nx = synth.make_clip(9, 272, 480, 1)
    residual = cv2.absdiff(nx[0], fr[0])
    diff = mfl.get_patch_diff(residual, 16)
    assert diff.dtype == np.float64 and np.array_equal(diff, F.patch_sums(residual))
    frag, positions = mfl.extract_important_patches(residual, diff)
    assert positions == F.topk_positions(diff) and np.array_equal(frag, F.gather_fragment(residual, positions))
    assert np.array_equal(mfl.get_original_frame_patches(fr[0], positions, 16, 224), F.gather_fragment(fr[0], positions))
    path, imp, pos2 = mfl.process_patches("a/b_1.png", "frame_diff", residual, 16, 224, 196)
    assert path == "a/b_1_residual_imp.png" and pos2 == positions
    assert np.array_equal(mfl.merge_fragments(frag, frag[::-1].copy()), cv2.addWeighted(frag, 0.5, frag[::-1].copy(), 0.5, 0))
    flow = mfl.calc_optical_flow_farneback(fr[0], nx[0])
    ref_flow = cv2.calcOpticalFlowFarneback(F.bgr2gray(fr[0]), F.bgr2gray(nx[0]), None, 0.5, 3, 15, 3, 5, 1.2, 0)
    assert np.abs(flow - ref_flow).max() < 2e-3
    from oracle import farneback as FB
    assert np.array_equal(mfl.flow_to_rgb(ref_flow), FB.flow_to_rgb(ref_flow))
    p = str(tmp_path / "vid_1.png")
    cv2.imwrite(p, fr[0])
    _, _, vec = mfl.get_deep_feature("resnet50", "vid", p, "original", "layer_stack")
    rows = mfl.process_video_feature([vec, vec], "resnet50", "layer_stack")
    assert rows.shape == (2, 13120)
    _, _, pv = mfl.get_deep_feature("resnet50", "vid", p, "original", "pool")
    assert mfl.process_video_feature([pv], "resnet50", "pool").shape == (1, 2051)
    _, _, vv = mfl.get_deep_feature("vit", "vid", p, "original", "pool")
    assert vv.shape == (2304,)
    with pytest.raises(ValueError):
        mfl.get_patch_diff(residual, 8)
    runtime._engine = None


def test_zero_pair_video_does_not_abort_batch(engine):
    """A video whose sampler produced no pairs gives NaN fragment blocks (imputed by the head), not an abort
    (the reference dies in np.concatenate)."""
    from relax_vqa_b200 import ops
    z = lambda w: torch.zeros((2, w), device="cuda")
    off_full = torch.tensor([0, 2], dtype=torch.int32, device="cuda")
    off_pair = torch.tensor([0, 0], dtype=torch.int32, device="cuda")
    feats = ops.temporal_mean_concat(z(13120), z(2304), z(13120), z(2051), z(2304), z(2304), off_full, off_pair)
    assert torch.isnan(feats[0, 15424:]).all() and not torch.isnan(feats[0, :15424]).any()
    score = ops.head_forward(engine.ctx, feats)
    assert torch.isfinite(score).all()


def test_dataset_driver_npy_and_mat_formats(engine, tmp_path):
    """8(f) rows 2-3: metadata CSV -> per-video .npy with the reference's names/shapes -> (N, D) .mat; resume."""
    import pandas as pd
    import scipy.io
    from relax_vqa_b200 import dataset_driver as dd
    from relax_vqa_b200.data_processing import extract_npy2mat as fmt
    frames_root, out_root = tmp_path / "frames", tmp_path / "out"
    vids = [("a1", 144, 256, 2), ("b2", 272, 480, 3), ("missing", 144, 256, 0)]
    for i, (vid, h, w, t) in enumerate(vids):
        d = frames_root / f"video_{i + 1}"
        d.mkdir(parents=True)
        if t:
            fr, nx = synth.make_clip(40 + i, h, w, t)
            for k in range(t):
                cv2.imwrite(str(d / f"{vid}_{k + 1}.png"), fr[k])
                cv2.imwrite(str(d / f"{vid}_{k + 1}_next.png"), nx[k])
    csv = tmp_path / "meta.csv"
    pd.DataFrame(dict(vid=[v[0] for v in vids], width=[v[2] for v in vids], height=[v[1] for v in vids], framerate=[29.97] * 3,
                      nb_frames=[60] * 3)).to_csv(csv, index=False)
    st = dd.run(str(csv), str(frames_root), str(out_root), "konvid_1k", engine=engine, batch_videos=2)
    assert [s for _, s in st][:2] == ["done", "done"] and st[2][1].startswith("error")
    p0 = dd.output_paths(str(out_root), "konvid_1k", 0)
    assert np.load(p0["full_resnet"]).shape == (2, 13120) and np.load(p0["frag_resnet"]).shape == (2, 15171)
    assert np.load(p0["full_vit"]).shape == (2, 2304) and np.load(p0["frag_vit"]).shape == (2, 4608)
    assert os.path.basename(p0["frag_resnet"]) == "video_1_resnet50_feature_map_original.npy"
    st2 = dd.run(str(csv), str(frames_root), str(out_root), "konvid_1k", engine=engine)
    assert [s for _, s in st2][:2] == ["skipped", "skipped"]
    d = fmt.features_dir(os.path.join(str(out_root), "features_merged_frag"), "resnet50", "layer_stack", "konvid_1k")
    mat = fmt.collate(d, 2, "resnet50")
    assert mat.shape == (2, 15171) and mat.dtype == np.float64
    assert np.allclose(mat[1], np.load(dd.output_paths(str(out_root), "konvid_1k", 1)["frag_resnet"]).mean(0))
    name = fmt.save_features(str(tmp_path / "mat"), "konvid_1k", mat, "resnet50")
    assert scipy.io.loadmat(name)["konvid_1k"].shape == (2, 15171)
    assert dd.frame_interval(29.97) == 14 and dd.frame_interval(1.5) == 1


@pytest.mark.parametrize("hw", [(100, 150), (333, 517)])
def test_odd_resolution_clip_vs_oracle(engine, hw):
    """LSVQ-style odd sizes: < 196 patches (zero-filled canvas), W % 16 != 0 (byte paths), shallow pyramid."""
    from relax_vqa_b200.engine import Clip
    fr, nx = synth.make_clip(77, hw[0], hw[1], 2)
    rsd, vsd = weights.seeded_resnet50_state_dict(1234), weights.seeded_vitb16_state_dict(4321)
    blocks = P.video_feature_blocks(fr, nx, rsd, vsd)
    ref = P.video_vector(blocks)
    clip = Clip(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda())
    inter = engine.fragments(clip.frames, clip.nexts, keep_intermediates=True)
    for t in range(2):
        assert np.array_equal(inter["diff_frag"][t].cpu().numpy(), blocks["pairs"][t]["diff_frag"])
        assert np.array_equal(inter["ori_frag"][t].cpu().numpy(), blocks["pairs"][t]["ori_frag"])
        assert np.abs(inter["flow"][t].cpu().numpy() - blocks["pairs"][t]["flow"]).max() < 2e-3
    got = engine.extract([clip])[0].cpu().numpy()
    e = seg_err(got, ref[None], VEC_SEGS)
    assert max(e) <= 1e-2, e


def test_srcc_and_mos_parity_over_several_videos(golden_dir):
    """north_star acceptance: |dMOS| <= 0.01 and SRCC >= 0.999 against the reference path's predictions, here the CPU
    oracle (itself pinned to the unmodified reference) on 8 distinct small clips with shared seeded weights."""
    import scipy.stats
    from relax_vqa_b200.engine import Clip, Engine
    s = np.load(os.path.join(golden_dir, "konvid_1k_scaler_imputer.npz"))
    rsd, vsd = weights.seeded_resnet50_state_dict(1234), weights.seeded_vitb16_state_dict(4321)
    hsd = weights.seeded_head_state_dict(99, swa_format=True)
    eng = Engine(0, rsd, vsd, hsd, s["imputer_mean"], s["scale"], s["minv"])
    clips, ref = [], []
    for i, hw in enumerate([(144, 256), (160, 240), (144, 256), (176, 208), (144, 256), (128, 320), (160, 240), (144, 256)]):
        fr, nx = synth.make_clip(500 + i, hw[0], hw[1], 2)
        fr = np.clip(fr.astype(np.int32) * (0.5 + 0.12 * i), 0, 255).astype(np.uint8)      # spread the content / scores
        nx = np.clip(nx.astype(np.int32) * (0.5 + 0.12 * i), 0, 255).astype(np.uint8)
        clips.append(Clip(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda()))
        vec = P.video_vector(P.video_feature_blocks(fr, nx, rsd, vsd))
        ref.append(P.predict(vec, weights.fix_state_dict(hsd), s["imputer_mean"], s["scale"], s["minv"], "konvid_1k"))
    _, score = eng.predict(clips, "konvid_1k")
    got = score.cpu().numpy()
    ref = np.array(ref)
    print("MOS gpu", np.round(got, 4), "oracle", np.round(ref, 4))
    assert np.abs(got - ref).max() <= 0.01
    assert scipy.stats.spearmanr(got, ref).correlation >= 0.999
    eng.close()


def test_srcc_540p_32_clips(golden_dir):
    """the SRCC / MOS acceptance on a BASELINE config's resolution - 32 distinct 540p clips (2 sampled
    pairs each), GPU vs the CPU oracle (cv2 Farneback, the reference's own flow) with shared seeded weights.  The head's
    BatchNorm statistics are calibrated on these clips (as training would) so that the scores spread over several MOS
    points and feature errors are amplified, not hidden behind a near-constant output.
    Segments that do not depend on the optical flow must agree within 1e-2 on every clip; the two merged-fragment blocks
    (flow-dependent top-196 selection: a tolerance-class stage) must agree within 1e-2 on every clip
    whose flow patch set equals cv2's, and the clips where it differs are counted and reported."""
    import cv2 as _cv2
    import scipy.stats
    from relax_vqa_b200.engine import Clip, Engine
    s = np.load(os.path.join(golden_dir, "konvid_1k_scaler_imputer.npz"))
    rsd, vsd = weights.seeded_resnet50_state_dict(1234), weights.seeded_vitb16_state_dict(4321)
    flow = lambda a, b: _cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    frames, vecs, ref_flow_pos = [], [], []
    for i in range(32):
        fr, nx = synth.make_clip(900 + i, 540, 960, 2)
        # diverse content so that the scores are well separated: contrast, brightness, blur, colour cast, inversion
        gain = 0.25 + 0.05 * i
        cast = np.array([1.0, 1.0 - 0.02 * (i % 7), 1.0 - 0.03 * (i % 5)], np.float32)
        def style(a):
            a = (a.astype(np.float32) - 128) * gain + 128 - 50 + 3.0 * i
            if i % 3 == 1:
                a = np.stack([_cv2.GaussianBlur(x, (0, 0), 1.0 + 0.15 * i) for x in a])
            if i % 4 == 3:
                a = 255.0 - a
            return np.clip(a * cast, 0, 255).astype(np.uint8)
        fr, nx = style(fr), style(nx)
        frames.append((fr, nx))
        blocks = P.video_feature_blocks(fr, nx, rsd, vsd, flow_fn=flow)
        vecs.append(P.video_vector(blocks))
        ref_flow_pos.append([set(map(tuple, p["flow_positions"])) for p in blocks["pairs"]])
    vecs = np.stack(vecs)
    # calibrated head: BN running statistics = statistics of fc1's outputs over these clips
    hsd = weights.seeded_head_state_dict(99)
    x = OH_impute(vecs, s)
    h = x.astype(np.float32) @ hsd["fc1.weight"].numpy().T + hsd["fc1.bias"].numpy()
    hsd["bn1.running_mean"] = torch.from_numpy(h.mean(0).astype(np.float32))
    hsd["bn1.running_var"] = torch.from_numpy(h.var(0).astype(np.float32) + 1e-6)
    hsd["fc3.weight"] = hsd["fc3.weight"] * 4.0                          # spread the outputs over several MOS points
    ref = np.array([P.predict(v, hsd, s["imputer_mean"], s["scale"], s["minv"], "konvid_1k") for v in vecs])
    eng = Engine(0, rsd, vsd, hsd, s["imputer_mean"], s["scale"], s["minv"])
    clips = [Clip(torch.from_numpy(fr).cuda(), torch.from_numpy(nx).cuda()) for fr, nx in frames]
    feats, score = eng.predict(clips, "konvid_1k")
    got, gf = score.cpu().numpy(), feats.cpu().numpy()
    # flow-dependent segments: frag_pool (2048 + 3 after the 15 frag_stack segments) and frag_vit_mer (last 3 x 768)
    n_seg = len(VEC_SEGS)
    flow_dep = set(range(15 + 3 + 15, 15 + 3 + 17)) | set(range(n_seg - 3, n_seg))
    worst_indep, worst_dep_same, differing = 0.0, 0.0, 0
    for i, c in enumerate(clips):
        e = seg_err(gf[i], vecs[i][None], VEC_SEGS)
        inter = eng.fragments(c.frames, c.nexts, keep_intermediates=True)
        same = all(set(map(tuple, inter["flow_positions"][t].cpu().numpy().tolist())) == ref_flow_pos[i][t] for t in range(2))
        worst_indep = max(worst_indep, max(v for k, v in enumerate(e) if k not in flow_dep))
        if same:
            worst_dep_same = max(worst_dep_same, max(e[k] for k in flow_dep))
        else:
            differing += 1
    srcc = scipy.stats.spearmanr(got, ref).correlation
    print("540p x32: max |dMOS|", np.abs(got - ref).max(), "SRCC", srcc, "score range", ref.min(), ref.max(), "worst seg err: flow-independent",
          worst_indep, "flow-dependent (same patch set)", worst_dep_same, "clips whose flow patch set differs from cv2's:", differing)
    _record("srcc_540p_32_clips", dict(max_abs_dmos=float(np.abs(got - ref).max()), srcc=float(srcc), worst_seg_err_flow_independent=float(worst_indep),
                                        worst_seg_err_flow_dependent_same_patches=float(worst_dep_same), clips_with_different_flow_patch_set=differing,
                                        score_min=float(ref.min()), score_max=float(ref.max())))
    gaps = np.diff(np.sort(ref))
    print("score gaps between neighbours: median", np.median(gaps), "min", gaps.min())
    assert ref.max() - ref.min() > 0.5 and np.median(gaps) > 5e-3      # a ranking test needs distinct, spread scores
    assert worst_indep <= 1e-2 and worst_dep_same <= 1e-2 and differing <= 8
    assert np.abs(got - ref).max() <= 0.01 and srcc >= 0.999
    eng.close()


def OH_impute(x, s):
    from oracle import head as OH
    return OH.impute_scale(x, s["imputer_mean"], s["scale"], s["minv"])
