"""Generates tests/golden/ref_train_synth.npz by running the UNMODIFIED reference's training code
(src/fine_tune.py::fine_tune_model with src/model_regression.py's Mlp and MAEAndRankLoss) on a small seeded problem.

Run in the build container only (needs /root/reference):   python tests/golden/gen_golden_train.py

Two runs, both with drop_rate = 0 (dropout masks come from torch's RNG and cannot be shared with another implementation):
  A  48 samples, batch 16 (the reference's loader is not shuffled), 8 epochs, no SWA   -> multi-batch steps, cosine LR, momentum,
     weight decay, BatchNorm running statistics
  B  full-batch (48), 8 epochs, SWA from epoch 6 -> AveragedModel averaging, the SWALR / cosine interplay, update_bn (one batch, so
     the reference's shuffled update_bn loader cannot change the statistics), the saved AveragedModel state-dict format
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import REF, install_shims  # noqa: E402

IN, HID, N = 96, 32, 48


def main():
    import torch
    scratch = tempfile.mkdtemp(prefix="relaxvqa_golden_train_")
    os.makedirs(os.path.join(scratch, "utils"))
    os.chdir(scratch)
    sys.path.insert(0, os.path.join(REF, "src"))
    install_shims()
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import fine_tune as ft
    from model_regression import Mlp
    ft.test_data_name = "golden"                     # a __main__ global the function reads when it saves (src/fine_tune.py:191)
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 1, (N, IN)).astype(np.float32)
    w = rng.standard_normal(IN).astype(np.float32)
    y = (50 + 12 * (X @ w) / np.sqrt(IN) + rng.standard_normal(N)).astype(np.float32)
    torch.manual_seed(11)
    init = Mlp(input_features=IN, hidden_features=HID, drop_rate=0.0)
    init_sd = {k: v.clone() for k, v in init.state_dict().items()}
    init_path = os.path.join(scratch, "init.pth")
    torch.save(init_sd, init_path)
    out = dict(X=X, y=y, IN=IN, HID=HID)
    for k, v in init_sd.items():
        out["init." + k] = v.numpy()
    hp = dict(loss_type="MAERankLoss", optimizer_type="sgd", initial_lr=0.05, weight_decay=0.005, l1_w=0.6, rank_w=1.0)
    for tag, batch, use_swa in (("A", 16, False), ("B", N, True)):
        model = Mlp(input_features=IN, hidden_features=HID, drop_rate=0.0)
        torch.manual_seed(3)
        res = ft.fine_tune_model(model, torch.device("cpu"), init_path, X, y, scratch, batch, 8, hp["loss_type"], hp["optimizer_type"],
                                 hp["initial_lr"], hp["weight_decay"], use_swa, hp["l1_w"], hp["rank_w"])
        res.eval()
        with torch.no_grad():
            pred = res(torch.from_numpy(X)).reshape(-1).numpy()
        for k, v in res.state_dict().items():
            out[f"{tag}.{k}"] = v.numpy()
        out[f"{tag}.pred"] = pred
        out[f"{tag}.batch"] = batch
    out.update({f"hp.{k}": v for k, v in hp.items() if isinstance(v, float)})
    np.savez_compressed(os.path.join(HERE, "ref_train_synth.npz"), **out)
    print("saved", sorted(k for k in out if k.startswith("B."))[:4], out["A.pred"][:3], out["B.pred"][:3])


if __name__ == "__main__":
    main()
