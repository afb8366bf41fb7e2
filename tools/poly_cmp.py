import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, cv2
from relax_vqa_b200 import ops, synth
from oracle import fragments as F
ctx = ops.Context(0)
d = "tests/golden/ref_example/original_5636101558"
a = cv2.imread(os.path.join(d, "5636101558_2.png")); b = cv2.imread(os.path.join(d, "5636101558_2_next.png"))
stored = cv2.imread(os.path.join(d, "5636101558_2_residual_of.png"))
r = ops.absdiff_patchsum(torch.from_numpy(a[None]).cuda(), torch.from_numpy(b[None]).cuda())
flow = ops.farneback(ctx, r["gray0"], r["gray1"])
rgb, sums, _ = ops.flow_to_rgb(flow)
diff = np.abs(rgb[0].cpu().numpy().astype(int) - stored.astype(int))
print("px differing", (diff != 0).any(-1).sum(), "frac", (diff != 0).any(-1).mean(), "max", diff.max())
ref = cv2.calcOpticalFlowFarneback(r["gray0"][0].cpu().numpy(), r["gray1"][0].cpu().numpy(), None, 0.5, 3, 15, 3, 5, 1.2, 0)
e = np.abs(flow[0].cpu().numpy() - ref); print("flow err vs cv2 max %.3e mean %.3e" % (e.max(), e.mean()))
np.save("gpurun_out/flow_%s.npy" % os.environ.get("TAG", "x"), flow[0].cpu().numpy())
