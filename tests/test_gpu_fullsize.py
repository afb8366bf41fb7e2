"""Full-size (BASELINE.json configs: 1080p, 2160p) checks through the C ABI: exact integer stages vs the oracle,
Farneback vs cv2 at 1080p, and size-independent properties at 2160p where the CPU oracle would take too long
(identical frames -> zero residual, translation of the pair -> translated sums, batch invariance,
host-buffer entry point == device-buffer entry point)."""
import cv2
import numpy as np
import pytest
import torch

from oracle import fragments as F
from relax_vqa_b200 import synth, weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from relax_vqa_b200.engine import Engine
    e = Engine(0, head_sd=weights.seeded_head_state_dict(), seed_if_missing=True)
    yield e
    e.close()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_1080p_pair_vs_oracle_and_cv2(engine):
    fr, nx = synth.make_clip(31, 1080, 1920, 1)
    out = engine.fragments(_dev(fr), _dev(nx), keep_intermediates=True)
    res = F.absdiff(nx[0], fr[0])
    frag, pos, sums = F.process_patches(res)
    assert sums.shape == (67, 120)
    assert np.array_equal(out["sums"][0].cpu().numpy().astype(np.float64), sums)
    assert [tuple(p) for p in out["positions"][0].cpu().numpy().tolist()] == pos
    assert np.array_equal(out["diff_frag"][0].cpu().numpy(), frag)
    assert np.array_equal(out["ori_frag"][0].cpu().numpy(), F.gather_fragment(fr[0], pos))
    ref = cv2.calcOpticalFlowFarneback(F.bgr2gray(fr[0]), F.bgr2gray(nx[0]), None, 0.5, 3, 15, 3, 5, 1.2, 0)
    err = np.abs(out["flow"][0].cpu().numpy() - ref)
    assert err.max() < 2e-3 and err.mean() < 2e-5, (err.max(), err.mean())


def test_2160p_properties(engine):
    from relax_vqa_b200 import ops
    H, W = 2160, 3840
    fr, nx = synth.make_clip(32, H, W, 1)
    f, n = _dev(fr), _dev(nx)
    # identical frames: zero residual sums and the first 196 raster cells selected (all-ties rule).  (The flow of
    # identical frames is NOT zero in OpenCV's Farneback: pixels whose displaced position is the last row / column
    # take the "outside" branch of UpdateMatrices - so the flow is compared with cv2 below instead.)
    same = engine.fragments(f, f, keep_intermediates=True)
    assert int(same["sums"].abs().sum()) == 0
    assert same["positions"][0].cpu().numpy().tolist() == [[i // 240, i % 240] for i in range(196)]
    # residual sums: checksum of checksums against numpy on the host (exact integers)
    r = ops.absdiff_patchsum(f, n)
    host = np.abs(nx[0].astype(np.int16) - fr[0].astype(np.int16))[:2160 - 2160 % 16].astype(np.int64)
    assert int(r["sums"].to(torch.int64).sum()) == int(host.sum())
    assert r["sums"].shape == (1, 135, 240)
    # translating both frames by a whole number of patches translates the selected positions
    out = engine.fragments(f, n, keep_intermediates=True)
    ref = cv2.calcOpticalFlowFarneback(F.bgr2gray(fr[0]), F.bgr2gray(nx[0]), None, 0.5, 3, 15, 3, 5, 1.2, 0)      # ~4 s on the host
    err = np.abs(out["flow"][0].cpu().numpy() - ref)
    assert err.max() < 2e-3 and err.mean() < 2e-5, (err.max(), err.mean())
    fs, ns = torch.roll(f, shifts=(32, 48), dims=(1, 2)), torch.roll(n, shifts=(32, 48), dims=(1, 2))
    out_s = engine.fragments(fs, ns, keep_intermediates=True)
    sums, sums_s = out["sums"][0], out_s["sums"][0]
    assert torch.equal(torch.roll(sums, shifts=(2, 3), dims=(0, 1)), sums_s)
    # batch invariance at full size: the pair alone == the pair inside a batch of two
    both = engine.fragments(torch.cat([f, fs]), torch.cat([n, ns]), keep_intermediates=True)
    assert torch.equal(both["merged_frag"][0], out["merged_frag"][0]) and torch.equal(both["flow"][1], out_s["flow"][0])


def test_host_entry_equals_device_entry(engine):
    from relax_vqa_b200.engine import Clip
    fr, nx = synth.make_clip(33, 540, 960, 3)
    clip = Clip(_dev(fr), _dev(nx))
    feats, score = engine.predict([clip, clip], "konvid_1k")
    host = [(torch.from_numpy(fr).pin_memory(), torch.from_numpy(nx).pin_memory())] * 2
    feats_h, score_h = engine.predict_host(host, "konvid_1k")
    assert torch.equal(feats, feats_h) and torch.equal(score.cpu(), score_h)
    assert torch.equal(feats[0], feats[1])
