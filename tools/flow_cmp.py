import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, cv2
from relax_vqa_b200 import ops, synth
from oracle import fragments as F
ctx = ops.Context(0)
for hw, seed in [((272, 480), 5), ((540, 960), 5), ((1080, 1920), 7)]:
    fr, nx = synth.make_clip(seed, hw[0], hw[1], 3)
    nx = np.array(nx); fr = np.array(fr)
    nx[2] = np.clip(fr[2].astype(np.int16) + np.random.default_rng(1).integers(-2, 3, fr[2].shape), 0, 255).astype(np.uint8)
    g0 = np.stack([F.bgr2gray(f) for f in fr]); g1 = np.stack([F.bgr2gray(f) for f in nx])
    refs = [cv2.calcOpticalFlowFarneback(g0[i], g1[i], None, 0.5, 3, 15, 3, 5, 1.2, 0) for i in range(3)]
    for impl in (2, 1, 0):
        ctx.set_flow_impl(impl)
        got = ops.farneback(ctx, torch.from_numpy(g0).cuda(), torch.from_numpy(g1).cuda()).cpu().numpy()
        one = ops.farneback(ctx, torch.from_numpy(g0[1:2]).cuda(), torch.from_numpy(g1[1:2]).cuda()).cpu().numpy()
        for i in range(3):
            err = np.abs(got[i] - refs[i]); p = np.unravel_index(err.argmax(), err.shape)
            print(hw, "impl", impl, "pair", i, "max %.2e mean %.2e at %s flow %s ref %s |ref|max %.2f" % (err.max(), err.mean(), p[:2], got[i][p[0], p[1]], refs[i][p[0], p[1]], np.abs(refs[i]).max()), "batch-inv", np.array_equal(one[0], got[1]), flush=True)
