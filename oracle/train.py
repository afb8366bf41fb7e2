"""Oracle (CPU, torch autograd) for the training half of the Mlp head.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Restates one optimisation step of src/model_regression.py:292-306 / src/fine_tune.py:158-166 (forward in train mode,
MAEAndRankLoss :61-89, backward, torch.optim.SGD with momentum and weight decay), AveragedModel.update_parameters (:400),
torch.optim.swa_utils.update_bn (:454-459) and the loop of fine_tune_model (fine_tune.py:130-193) with explicit dropout
keep-masks, so that a second implementation can be fed the same masks.  Pinned against the UNMODIFIED reference's
fine_tune_model (tests/golden/gen_golden_train.py -> ref_train_synth.npz; tests/test_oracle_train.py).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

PARAMS = ["fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"]


def mae_rank_loss(pred, y, l1_w, rank_w):
    """MAEAndRankLoss.forward with use_margin = False; pred, y: (B, 1)."""
    l_mae = F.l1_loss(pred, y, reduction="mean") * l1_w
    n = pred.size(0)
    pd = pred.unsqueeze(1) - pred.unsqueeze(0)
    td = y.unsqueeze(1) - y.unsqueeze(0)
    l_rank = F.relu(td - torch.sign(td) * pd).sum() / (n * (n - 1))
    return l_mae + l_rank * rank_w


def forward(p, run_mean, run_var, X, train, masks=None, drop_rate=0.0, bn_momentum=0.1):
    """Mlp.forward; train: batch statistics and in-place running-statistics update; masks: (m1, m2) keep masks or None."""
    h = F.linear(X, p["fc1.weight"], p["fc1.bias"])
    h = F.batch_norm(h, run_mean, run_var, p["bn1.weight"], p["bn1.bias"], training=train, momentum=bn_momentum, eps=1e-5)
    h = F.gelu(h)
    if train and masks is not None and drop_rate > 0:
        h = h * masks[0].float() / (1.0 - drop_rate)
    h = F.gelu(F.linear(h, p["fc2.weight"], p["fc2.bias"]))
    if train and masks is not None and drop_rate > 0:
        h = h * masks[1].float() / (1.0 - drop_rate)
    return F.linear(h, p["fc3.weight"], p["fc3.bias"])


class Trainer:
    def __init__(self, state_dict, drop_rate=0.0):
        self.p = OrderedDict((k, state_dict[k].detach().clone().float().requires_grad_(True)) for k in PARAMS)
        self.run_mean = state_dict["bn1.running_mean"].detach().clone().float()
        self.run_var = state_dict["bn1.running_var"].detach().clone().float()
        self.bufs, self.drop_rate = None, drop_rate
        self.swa = OrderedDict((k, v.detach().clone()) for k, v in self.p.items())      # AveragedModel(model): a deep copy
        self.swa_mean, self.swa_var, self.n_averaged = self.run_mean.clone(), self.run_var.clone(), 0

    def step(self, X, y, lr, momentum, weight_decay, l1_w, rank_w, masks=None):
        X, y = torch.as_tensor(X).float(), torch.as_tensor(y).float().view(-1, 1)
        for v in self.p.values():
            v.grad = None
        pred = forward(self.p, self.run_mean, self.run_var, X, True, masks, self.drop_rate)
        loss = mae_rank_loss(pred, y, l1_w, rank_w)
        loss.backward()
        with torch.no_grad():                       # torch.optim.SGD
            first = self.bufs is None
            if first:
                self.bufs = {}
            for k, v in self.p.items():
                g = v.grad + weight_decay * v
                self.bufs[k] = g.clone() if first else self.bufs[k] * momentum + g
                v -= lr * self.bufs[k]
        return float(loss)

    def swa_update(self):
        with torch.no_grad():
            for k in self.p:
                if self.n_averaged == 0:
                    self.swa[k].copy_(self.p[k])
                else:
                    self.swa[k] += (self.p[k] - self.swa[k]) / (self.n_averaged + 1)
        self.n_averaged += 1

    def update_bn(self, batches):
        """torch.optim.swa_utils.update_bn on the SWA model: reset, momentum = None (cumulative average)."""
        self.swa_mean.zero_(); self.swa_var.fill_(1.0)
        with torch.no_grad():
            for i, xb in enumerate(batches):
                forward(self.swa, self.swa_mean, self.swa_var, torch.as_tensor(xb).float(), True, bn_momentum=1.0 / (i + 1))

    @torch.no_grad()
    def predict(self, X, swa=False):
        p, m, v = (self.swa, self.swa_mean, self.swa_var) if swa else (self.p, self.run_mean, self.run_var)
        return forward(p, m, v, torch.as_tensor(X).float(), False).reshape(-1).numpy()

    def state_dict(self, swa=False):
        p, m, v = (self.swa, self.swa_mean, self.swa_var) if swa else (self.p, self.run_mean, self.run_var)
        sd = OrderedDict((k, t.detach().clone()) for k, t in p.items())
        sd["bn1.running_mean"], sd["bn1.running_var"] = m.clone(), v.clone()
        return sd


def fine_tune(init_sd, X, y, batch_size, epochs, initial_lr, weight_decay, use_swa, l1_w, rank_w, drop_rate=0.0, masks_fn=None,
              update_bn_perm=None):
    """The loop of fine_tune_model (src/fine_tune.py:130-193).  -> (Trainer, swa_is_final, per-epoch losses)."""
    from torch.optim.lr_scheduler import CosineAnnealingLR
    from torch.optim.swa_utils import SWALR
    tr = Trainer(init_sd, drop_rate)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=initial_lr, momentum=0.9, weight_decay=weight_decay)
    sched = CosineAnnealingLR(opt, T_max=epochs, eta_min=1e-5)
    swa_sched = SWALR(opt, swa_lr=initial_lr, anneal_strategy="cos") if use_swa else None
    swa_start = int(epochs * 0.75) if use_swa else epochs
    X, y = np.asarray(X, np.float32), np.asarray(y, np.float32)
    losses, step_no = [], 0
    for epoch in range(epochs):
        tot = 0.0
        for i in range(0, len(X), batch_size):
            xb, yb = X[i:i + batch_size], y[i:i + batch_size]
            masks = masks_fn(step_no, len(xb)) if masks_fn else None
            tot += tr.step(xb, yb, opt.param_groups[0]["lr"], 0.9, weight_decay, l1_w, rank_w, masks) * len(xb)
            step_no += 1
            opt.step()
        losses.append(tot / len(X))
        sched.step()
        if use_swa and epoch >= swa_start:
            tr.swa_update()
            swa_sched.step()
    swa_final = use_swa and epochs - 1 >= swa_start
    if swa_final:
        perm = np.arange(len(X)) if update_bn_perm is None else np.asarray(update_bn_perm)
        tr.update_bn(X[perm[i:i + batch_size]] for i in range(0, len(X), batch_size))
    return tr, swa_final, losses
