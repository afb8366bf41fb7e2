"""Drop-in for the training half of src/fine_tune.py: ``fine_tune_model`` (ref :130-193) with every optimisation step on
the GPU (libb200vqa nn_train.cu through model_regression.HeadTrainer)."""
import os

import numpy as np
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR
from torch.optim.swa_utils import SWALR

from .model_regression import HeadTrainer, MAEAndRankLoss, make_optimizer
from .weights import fix_state_dict  # noqa: F401  (ref :99-109)


def fine_tune_model(model, device, model_path, X_fine_tune, y_fine_tune, save_path, batch_size, epochs, loss_type, optimizer_type,
                    initial_lr, weight_decay, use_swa, l1_w, rank_w, test_data_name="fine_tune", update_bn_perm=None):
    """ref :130-193.  model: a HeadTrainer (its drop_rate is the Mlp's); model_path: the pre-trained head to start from
    (a .pth path or a state dict).  The loader is not shuffled (ref :141); SWA starts at 75 % of the epochs; after training
    the SWA model's BatchNorm statistics are recomputed over the data (update_bn; ``update_bn_perm`` fixes the order the
    reference's shuffled loader would draw) and the result is saved in AveragedModel format.  Returns the state dict saved."""
    if loss_type != 'MAERankLoss' or optimizer_type != 'sgd':
        raise ValueError("the device trainer implements loss_type 'MAERankLoss' with optimizer_type 'sgd' (the reference's defaults)")
    sd = torch.load(model_path, map_location='cpu') if isinstance(model_path, (str, os.PathLike)) else model_path
    model.load_state_dict(sd)
    X = model._dev(np.asarray(X_fine_tune, np.float32))
    y = model._dev(np.asarray(y_fine_tune, np.float32))
    criterion = MAEAndRankLoss(l1_w, rank_w)
    optimizer = make_optimizer(initial_lr, weight_decay)
    scheduler = CosineAnnealingLR(optimizer, T_max=epochs, eta_min=1e-5)
    swa_scheduler = SWALR(optimizer, swa_lr=initial_lr, anneal_strategy='cos') if use_swa else None
    swa_start = int(epochs * 0.75) if use_swa else epochs
    losses, epoch = [], -1
    for epoch in range(epochs):
        g = optimizer.param_groups[0]
        total = torch.zeros((), device=model.device)
        for i in range(0, X.shape[0], batch_size):
            xb, yb = X[i:i + batch_size], y[i:i + batch_size]
            total += model.step(xb, yb, g["lr"], g["momentum"], g["weight_decay"], l1_w, rank_w) * xb.shape[0]
            optimizer.step()             # no-op on the dummy parameter; keeps torch's scheduler bookkeeping in order
        losses.append(float(total.item()) / X.shape[0])
        scheduler.step()
        if use_swa and epoch >= swa_start:
            model.swa_update()
            swa_scheduler.step()
    swa_final = use_swa and epoch >= swa_start
    if swa_final:
        perm = torch.arange(X.shape[0]) if update_bn_perm is None else torch.as_tensor(update_bn_perm)
        perm = perm.to(model.device)
        model.update_bn((X[perm[i:i + batch_size]] for i in range(0, X.shape[0], batch_size)), swa=True)
    out = model.state_dict(swa=swa_final)
    if save_path:
        os.makedirs(save_path, exist_ok=True)
        torch.save(out, os.path.join(save_path, f"{test_data_name}_relaxvqa_fine_tuned_model.pth"))
    model.fine_tune_losses = losses
    return out
