"""GPU parity of the head trainer through the C ABI: against the UNMODIFIED reference's
fine_tune_model (tests/golden/ref_train_synth.npz) and, at the real head size with dropout, against the CPU oracle fed the
same keep-masks.  fp32 arithmetic on both sides: parameters within 1e-4 (relative to the tensor's max), losses within 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import train as OT
from relax_vqa_b200 import weights

pytestmark = pytest.mark.gpu

KEYS = ["fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "bn1.running_mean", "bn1.running_var", "fc2.weight", "fc2.bias",
        "fc3.weight", "fc3.bias"]


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-6)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_fine_tune_model_vs_reference_golden(golden_dir, tmp_path, tag):
    """src/fine_tune.py:130-193 run by the unmodified reference (drop_rate 0): A = 3 batches per epoch, no SWA; B = full batch,
    SWA from epoch 6, update_bn, AveragedModel file format."""
    from relax_vqa_b200.fine_tune import fine_tune_model
    from relax_vqa_b200.model_regression import HeadTrainer
    g = np.load(os.path.join(golden_dir, "ref_train_synth.npz"))
    init = {k: torch.from_numpy(g["init." + k]) for k in KEYS}
    init_path = str(tmp_path / "init.pth")
    torch.save(init, init_path)
    use_swa = tag == "B"
    model = HeadTrainer(int(g["IN"]), int(g["HID"]), drop_rate=0.0)
    sd = fine_tune_model(model, "cuda", init_path, g["X"], g["y"], str(tmp_path), int(g[f"{tag}.batch"]), 8, "MAERankLoss", "sgd",
                         float(g["hp.initial_lr"]), float(g["hp.weight_decay"]), use_swa, float(g["hp.l1_w"]), float(g["hp.rank_w"]),
                         test_data_name="golden")
    prefix = f"{tag}.module." if use_swa else f"{tag}."
    for k in KEYS:
        e = rel(sd[("module." if use_swa else "") + k].numpy(), g[prefix + k])
        assert e < 1e-4, (k, e)
    if use_swa:
        assert int(sd["n_averaged"]) == int(g["B.n_averaged"]) == 2
        assert list(sd)[0] == "n_averaged" and all(k.startswith("module.") for k in list(sd)[1:])
    pred = model.predict(g["X"], swa=use_swa).cpu().numpy()
    assert rel(pred, g[f"{tag}.pred"]) < 1e-3
    saved = torch.load(os.path.join(str(tmp_path), "golden_relaxvqa_fine_tuned_model.pth"))
    assert weights.fix_state_dict(saved)["fc1.weight"].shape == (int(g["HID"]), int(g["IN"]))
    assert model.fine_tune_losses[-1] < model.fine_tune_losses[0]
    model.close()


def test_steps_with_dropout_at_the_real_head_size_vs_oracle():
    """Mlp(35203, 256), batch 48, drop_rate 0.2 with shared keep-masks: 4 SGD steps, then SWA and a 2-batch update_bn."""
    from relax_vqa_b200.model_regression import HeadTrainer, init_state_dict
    torch.manual_seed(7)
    init = init_state_dict(35203, 256)
    rng = np.random.default_rng(2)
    X = rng.uniform(0, 1, (96, 35203)).astype(np.float32)
    y = rng.uniform(20, 80, 96).astype(np.float32)
    masks = [(rng.uniform(size=(48, 256)) >= 0.2, rng.uniform(size=(48, 128)) >= 0.2) for _ in range(4)]
    ref = OT.Trainer(init, drop_rate=0.2)
    model = HeadTrainer(35203, 256, drop_rate=0.2, state_dict=init)
    for s in range(4):
        xb, yb = X[48 * (s % 2):48 * (s % 2) + 48], y[48 * (s % 2):48 * (s % 2) + 48]
        m = (torch.from_numpy(masks[s][0]), torch.from_numpy(masks[s][1]))
        l_ref = ref.step(xb, yb, 0.05, 0.9, 0.005, 0.6, 1.0, m)
        l_gpu = float(model.step(xb, yb, 0.05, 0.9, 0.005, 0.6, 1.0, masks=m))
        assert abs(l_gpu - l_ref) < 1e-4 * max(1.0, abs(l_ref)), (s, l_gpu, l_ref)
        if s >= 2:
            ref.swa_update(); model.swa_update()
    ref.update_bn([X[:48], X[48:]]); model.update_bn([X[:48], X[48:]], swa=True)
    for swa in (False, True):
        got, want = model.state_dict(swa=False) if not swa else weights.fix_state_dict(model.state_dict(swa=True)), ref.state_dict(swa=swa)
        for k in KEYS:
            e = rel(got[k].numpy(), want[k].numpy())
            assert e < 1e-4, (swa, k, e)
        assert rel(model.predict(X, swa=swa).cpu().numpy(), ref.predict(X, swa=swa)) < 1e-3
    # the trained head drops into the inference path (engine head kernels) unchanged
    from relax_vqa_b200 import ops
    ctx = ops.Context(0)
    ops.load_head(ctx, model.state_dict(swa=True), np.zeros(35203), np.ones(35203), np.zeros(35203))
    score = ops.head_forward(ctx, torch.from_numpy(X[:5]).cuda()).cpu().numpy()
    assert rel(score, ref.predict(X[:5], swa=True)) < 1e-3
    ctx.close(); model.close()


def test_train_and_evaluate_kfold_swa(tmp_path):
    """ref :335-471 end to end on a small learnable problem: k-fold, cosine LR, SWA, selection by RMSE; the returned state dict
    predicts better than the initial model and has a loadable format."""
    from relax_vqa_b200.model_regression import HeadTrainer, compute_correlation_metrics, train_and_evaluate
    rng = np.random.default_rng(4)
    X = rng.uniform(0, 1, (150, 64)).astype(np.float32)
    w = rng.standard_normal(64).astype(np.float32)
    y = (3.0 + (X @ w) / 4.0 + 0.05 * rng.standard_normal(150)).astype(np.float32)
    cfg = dict(n_repeats=1, n_splits=3, batch_size=32, epochs=30, hidden_features=32, drop_rate=0.1, loss_type="MAERankLoss",
               optimizer_type="sgd", select_criteria="byrmse", initial_lr=0.05, weight_decay=0.0005, patience=5, l1_w=0.6, rank_w=1.0,
               use_swa=True, seed=1)
    torch.manual_seed(0)
    sd, tl, vl = train_and_evaluate(X, y, cfg)
    assert len(tl) == 3 and all(l[-1] < l[0] for l in tl) and len(vl) == 3
    model = HeadTrainer(64, 32, drop_rate=0.0, state_dict=sd)
    pred = model.predict(X).cpu().numpy()
    _, plcc, rmse, srcc, krcc = compute_correlation_metrics(y, pred)
    print("train_and_evaluate: SRCC", srcc, "PLCC", plcc, "RMSE", rmse)
    assert srcc > 0.8 and rmse < 0.5 * np.std(y) * 2
    model.close()
