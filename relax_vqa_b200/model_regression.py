"""Inference half of src/model_regression.py: ``Mlp`` with the reference's constructor and state-dict
format, evaluated by the fused head kernels (imputer + scaler + fc1 + BN + GELU + fc2 + GELU + fc3)."""
import numpy as np
import torch

from . import ops, runtime, weights


class Mlp:
    """ref :37-58.  Inference only (eval mode): training stays with the reference's PyTorch module."""

    def __init__(self, input_features, hidden_features=256, out_features=1, drop_rate=0.2, act_layer=None):
        if input_features != weights.FEATURE_DIM or hidden_features != 256 or out_features != 1:
            raise ValueError("the fused head implements Mlp(35203, 256, 1) (src/demo_test.py:183-185)")
        self._sd = None
        self._pre = (np.zeros(input_features), np.ones(input_features), np.zeros(input_features))

    def set_preprocessing(self, imputer=None, scaler=None):
        """Fitted SimpleImputer / MinMaxScaler objects (or None): uses statistics_, scale_, min_ directly,
        which also sidesteps the sklearn-1.3 pickle incompatibility noted in SURVEY.md 0.4."""
        n = weights.FEATURE_DIM
        self._pre = (np.asarray(imputer.statistics_, np.float64) if imputer is not None else np.zeros(n),
                     np.asarray(scaler.scale_, np.float64) if scaler is not None else np.ones(n),
                     np.asarray(scaler.min_, np.float64) if scaler is not None else np.zeros(n))
        if self._sd is not None:
            self._install()

    def load_state_dict(self, state_dict):
        self._sd = weights.fix_state_dict(state_dict)
        missing = [k for k, _ in weights.head_spec() if k not in self._sd]
        if missing:
            raise RuntimeError(f"Missing key(s) in state_dict: {missing}")
        self._install()

    def _install(self):
        runtime.engine().load_head(self._sd, *self._pre)

    def eval(self):
        return self

    def to(self, device):
        return self

    def __call__(self, features):
        """features: (V, 35203) raw (un-imputed, un-scaled) float array or tensor -> (V, 1) tensor."""
        if self._sd is None:
            raise RuntimeError("load_state_dict first")
        x = torch.as_tensor(np.asarray(features, dtype=np.float32) if not torch.is_tensor(features) else features)
        x = x.to(runtime.engine().device, torch.float32).contiguous().reshape(-1, weights.FEATURE_DIM)
        return ops.head_forward(runtime.engine().ctx, x).reshape(-1, 1)


# ======================================================================================================================
# Training half: the arithmetic of every optimisation step runs in libb200vqa (nn_train.cu);
# this module keeps the reference's schedule and bookkeeping (src/model_regression.py:292-471).
# ======================================================================================================================
import copy
import ctypes as C
from collections import OrderedDict

from . import _lib

_PARAM_KEYS = ["fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "bn1.running_mean", "bn1.running_var",
               "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"]


class MAEAndRankLoss:
    """ref :61-89 (use_margin is not supported on the device path: the reference never enables it).  The device step
    computes this loss and its gradient itself; calling the object evaluates it for reporting (validation loss)."""

    def __init__(self, l1_w=1.0, rank_w=1.0, margin=0.0, use_margin=False):
        if use_margin and margin > 0:
            raise ValueError("MAEAndRankLoss with a margin is not implemented on the device path")
        self.l1_w, self.rank_w, self.margin, self.use_margin = l1_w, rank_w, margin, use_margin

    def __call__(self, y_pred, y_true):
        p = torch.as_tensor(y_pred, dtype=torch.float32).reshape(-1).cpu()
        y = torch.as_tensor(y_true, dtype=torch.float32).reshape(-1).cpu()
        n = p.numel()
        l_mae = (p - y).abs().mean() * self.l1_w
        pd, td = p[:, None] - p[None, :], y[:, None] - y[None, :]
        l_rank = torch.relu(td - torch.sign(td) * pd).sum() / (n * (n - 1)) if n > 1 else torch.tensor(float("nan"))
        return l_mae + l_rank * self.rank_w


def init_state_dict(input_features, hidden_features=256, out_features=1):
    """PyTorch's default initialisation of the reference's Mlp, drawn from torch's global CPU generator in the order the
    reference constructs the layers (fc1, bn1, fc2, fc3; ref :40-47): with the same torch.manual_seed the head starts
    from the same parameters as the reference's ``Mlp(...)``.  Host-side, once per model."""
    fc1 = torch.nn.Linear(input_features, hidden_features)
    bn1 = torch.nn.BatchNorm1d(hidden_features)
    fc2 = torch.nn.Linear(hidden_features, hidden_features // 2)
    fc3 = torch.nn.Linear(hidden_features // 2, out_features)
    sd = OrderedDict()
    for name, m in (("fc1", fc1), ("bn1", bn1), ("fc2", fc2), ("fc3", fc3)):
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


class HeadTrainer:
    """Device-side trainer of ``Mlp(input_features, hidden_features)``: parameters, gradients, momentum and the SWA average
    live in HBM; one ``step`` = forward (train mode) + MAE/rank loss + backward + SGD update."""

    def __init__(self, input_features, hidden_features=256, drop_rate=0.2, device=None, state_dict=None, seed=None):
        from . import ops
        self.lib = _lib.load()
        self.in_features, self.hidden, self.drop_rate = int(input_features), int(hidden_features), float(drop_rate)
        if device is None and runtime._engine is not None:
            self._ctx, self._own_ctx = runtime.engine().ctx, False
        else:
            self._ctx, self._own_ctx = ops.Context(0 if device is None else torch.device(device).index or 0), True
        self.device = self._ctx.device
        h = C.c_void_p()
        _lib.check(self.lib.b200vqa_trainer_create(self._ctx.h, self.in_features, self.hidden, C.byref(h)), "trainer_create")
        self.h = h
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(0 if seed is None else int(seed))
        self._loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.load_state_dict(state_dict if state_dict is not None else init_state_dict(input_features, hidden_features))

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200vqa_trainer_destroy(self.h)
            self.h = None
        if getattr(self, "_own_ctx", False):
            self._ctx.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ------------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """Also resets the momentum buffers and re-creates the SWA copy (AveragedModel(model), ref :388)."""
        sd = weights.fix_state_dict(state_dict)
        keep = [sd[k].detach().to(torch.float32).contiguous().cpu() for k in _PARAM_KEYS]
        if keep[0].shape != (self.hidden, self.in_features):
            raise RuntimeError(f"size mismatch for fc1.weight: {tuple(keep[0].shape)}")
        _lib.check(self.lib.b200vqa_trainer_set_params(self.h, *[C.c_void_p(t.data_ptr()) for t in keep]), "trainer_set_params")
        self.num_batches_tracked = int(sd.get("bn1.num_batches_tracked", torch.tensor(0)))
        self.n_averaged = 0

    def state_dict(self, swa=False):
        """Mlp state dict; swa=True: the AveragedModel format the reference saves (``module.`` prefix + ``n_averaged``,
        ref :388, :715), which fix_state_dict / demo_test load."""
        out = [torch.empty(s, dtype=torch.float32) for s in
               ((self.hidden, self.in_features), (self.hidden,), (self.hidden,), (self.hidden,), (self.hidden,), (self.hidden,),
                (self.hidden // 2, self.hidden), (self.hidden // 2,), (1, self.hidden // 2), (1,))]
        _lib.check(self.lib.b200vqa_trainer_get_params(self.h, int(bool(swa)), *[C.c_void_p(t.data_ptr()) for t in out]), "trainer_get_params")
        sd = OrderedDict(zip(_PARAM_KEYS, out))
        sd["bn1.num_batches_tracked"] = torch.tensor(self.num_batches_tracked, dtype=torch.long)
        sd = OrderedDict((k, sd[k]) for k in [k for k, _ in weights.head_spec(self.in_features, self.hidden)])
        if not swa:
            return sd
        wrapped = OrderedDict(n_averaged=torch.tensor(self.n_averaged, dtype=torch.long))
        for k, v in sd.items():
            wrapped["module." + k] = v
        return wrapped

    # ---- one optimisation step -------------------------------------------------------------------------------------
    def _dev(self, a):
        t = torch.as_tensor(a) if not torch.is_tensor(a) else a
        return t.to(self.device, torch.float32).contiguous()

    def step(self, X, y, lr, momentum=0.9, weight_decay=0.0, l1_w=1.0, rank_w=1.0, masks=None):
        """X [B, in], y [B] -> loss (0-dim device tensor).  Dropout keep-masks are drawn on the device from this
        trainer's generator unless given (masks = (m1 [B, hidden], m2 [B, hidden/2]) uint8)."""
        X, y = self._dev(X), self._dev(y).reshape(-1)
        B = X.shape[0]
        m1 = m2 = None
        if self.drop_rate > 0:
            if masks is None:
                m1 = (torch.rand((B, self.hidden), device=self.device, generator=self._gen) >= self.drop_rate).to(torch.uint8)
                m2 = (torch.rand((B, self.hidden // 2), device=self.device, generator=self._gen) >= self.drop_rate).to(torch.uint8)
            else:
                m1, m2 = (torch.as_tensor(m).to(self.device, torch.uint8).contiguous() for m in masks)
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.b200vqa_trainer_step(self.h, _lib.ptr(X), _lib.ptr(y), B, _lib.ptr(m1), _lib.ptr(m2), self.drop_rate, float(lr),
                                                 float(momentum), float(weight_decay), float(l1_w), float(rank_w), _lib.ptr(loss),
                                                 _lib.stream_ptr(self.device)), "trainer_step")
        self.num_batches_tracked += 1
        return loss

    def swa_update(self):
        """AveragedModel.update_parameters(model) (ref :400)."""
        _lib.check(self.lib.b200vqa_trainer_swa_update(self.h, _lib.stream_ptr(self.device)), "trainer_swa_update")
        self.n_averaged += 1

    def predict(self, X, swa=False, batch_size=4096):
        """Eval-mode forward -> (N,) device tensor."""
        X = self._dev(X)
        out = torch.empty((X.shape[0],), dtype=torch.float32, device=self.device)
        for i in range(0, X.shape[0], batch_size):
            xb = X[i:i + batch_size].contiguous()
            _lib.check(self.lib.b200vqa_trainer_predict(self.h, int(bool(swa)), _lib.ptr(xb), xb.shape[0], _lib.ptr(out[i:i + batch_size]),
                                                        _lib.stream_ptr(self.device)), "trainer_predict")
        return out

    def update_bn(self, batches, swa=True):
        """torch.optim.swa_utils.update_bn(loader, model) (ref :454-459): cumulative average of the batch statistics."""
        for i, xb in enumerate(batches):
            xb = self._dev(xb)
            _lib.check(self.lib.b200vqa_trainer_update_bn(self.h, int(bool(swa)), _lib.ptr(xb), xb.shape[0], i, _lib.stream_ptr(self.device)),
                       "trainer_update_bn")


# ---- the reference's loop, same names -------------------------------------------------------------------------------
def train_one_epoch(model, train_loader, criterion, optimizer):
    """ref :292-306.  model: HeadTrainer; train_loader: iterable of (inputs, targets); criterion: MAEAndRankLoss (weights only);
    optimizer: a torch optimizer over a dummy parameter that only carries the hyper-parameters the schedulers edit
    (lr, momentum, weight_decay) - see make_optimizer."""
    g = optimizer.param_groups[0]
    total, n = torch.zeros((), device=model.device), 0
    for inputs, targets in train_loader:
        loss = model.step(inputs, targets, g["lr"], g.get("momentum", 0.0), g.get("weight_decay", 0.0), criterion.l1_w, criterion.rank_w)
        optimizer.step()                 # no-op on the dummy parameter; keeps torch's scheduler bookkeeping in order
        total += loss * inputs.shape[0]
        n += inputs.shape[0]
    return float(total.item()) / max(n, 1)


def evaluate(model, val_loader, criterion, swa=False):
    """ref :308-322 -> (validation loss, predictions)."""
    val_loss, preds, n = 0.0, [], 0
    for inputs, targets in val_loader:
        out = model.predict(inputs, swa=swa).cpu()
        preds.extend(out.tolist())
        val_loss += float(criterion(out, targets)) * inputs.shape[0]
        n += inputs.shape[0]
    return val_loss / max(n, 1), np.array(preds)


def make_optimizer(initial_lr, weight_decay, momentum=0.9):
    """torch.optim.SGD over a dummy parameter: the reference's LR schedulers (CosineAnnealingLR, SWALR; ref :376-390) are
    driven unchanged and their lr is read back each epoch - the schedule is host control flow, the update runs on the GPU."""
    dummy = torch.nn.Parameter(torch.zeros(1))
    return torch.optim.SGD([dummy], lr=initial_lr, momentum=momentum, weight_decay=weight_decay)


def logistic_func(X, bayta1, bayta2, bayta3, bayta4):
    """ref :137-140."""
    return bayta2 + (bayta1 - bayta2) / (1 + np.exp(-(X - bayta3) / np.abs(bayta4)))


def compute_correlation_metrics(y_true, y_pred):
    """ref :142-161 -> (y_pred_logistic, plcc, rmse, srcc, krcc); host-side metric code (scipy)."""
    import scipy.optimize
    import scipy.stats
    y_true, y_pred = np.asarray(y_true, float), np.asarray(y_pred, float)
    beta = [np.max(y_true), np.min(y_true), np.mean(y_pred), 0.5]
    popt, _ = scipy.optimize.curve_fit(logistic_func, y_pred, y_true, p0=beta, maxfev=100000000)
    yl = logistic_func(y_pred, *popt)
    plcc = scipy.stats.pearsonr(y_true, yl)[0]
    rmse = float(np.sqrt(np.mean((y_true - yl) ** 2)))
    return yl, plcc, rmse, scipy.stats.spearmanr(y_true, y_pred)[0], scipy.stats.kendalltau(y_true, y_pred)[0]


def _batches(X, y, batch_size, perm=None):
    idx = np.arange(len(X)) if perm is None else perm
    for i in range(0, len(idx), batch_size):
        j = idx[i:i + batch_size]
        yield X[j], y[j]


def train_and_evaluate(X_train, y_train, config):
    """ref :335-471: k-fold training with SGD + cosine LR, optional SWA from 70 % of the epochs, model selection by RMSE or
    KRCC on the validation fold, early stopping once SWA is active.  Returns (best_state_dict, all_train_losses,
    all_val_losses); best_state_dict is in AveragedModel format when the selected model is the SWA average."""
    from sklearn.model_selection import KFold
    from torch.optim.lr_scheduler import CosineAnnealingLR
    from torch.optim.swa_utils import SWALR
    n_splits, batch_size, epochs = config['n_splits'], config['batch_size'], config['epochs']
    hidden, drop_rate = config['hidden_features'], config['drop_rate']
    select_criteria, initial_lr, weight_decay = config['select_criteria'], config['initial_lr'], config['weight_decay']
    patience, use_swa = config['patience'], config.get('use_swa', False)
    if config.get('optimizer_type', 'sgd') != 'sgd' or config.get('loss_type', 'MAERankLoss') != 'MAERankLoss':
        raise ValueError("the device trainer implements the reference's defaults: optimizer_type 'sgd', loss_type 'MAERankLoss'")
    X_train, y_train = np.asarray(X_train, np.float32), np.asarray(y_train, np.float32)
    rng = np.random.default_rng(config.get('seed', 0))
    kf = KFold(n_splits=n_splits, shuffle=True, random_state=42)
    best_sd, best_metric = None, (float('inf') if select_criteria == 'byrmse' else float('-inf'))
    all_train_losses, all_val_losses = [], []
    for fold, (train_idx, val_idx) in enumerate(kf.split(X_train)):
        Xtr, Xva, ytr, yva = X_train[train_idx], X_train[val_idx], y_train[train_idx], y_train[val_idx]
        model = HeadTrainer(Xtr.shape[1], hidden, drop_rate, seed=config.get('seed', 0) + fold)
        Xtr_d, ytr_d = model._dev(Xtr), model._dev(ytr)
        Xva_d, yva_t = model._dev(Xva), torch.from_numpy(yva)
        criterion = MAEAndRankLoss(config['l1_w'], config['rank_w'])
        optimizer = make_optimizer(initial_lr, weight_decay)
        scheduler = CosineAnnealingLR(optimizer, T_max=epochs, eta_min=1e-5)
        swa_scheduler = SWALR(optimizer, swa_lr=initial_lr, anneal_strategy='cos') if use_swa else None
        swa_start = int(epochs * 0.7) if use_swa else epochs
        train_losses, val_losses = [], []
        best_val_loss, epochs_no_improve, early_stop_active, fold_best = float('inf'), 0, False, None
        for epoch in range(epochs):
            perm = torch.from_numpy(rng.permutation(len(Xtr))).to(model.device)           # DataLoader(shuffle=True)
            loader = ((Xtr_d[perm[i:i + batch_size]], ytr_d[perm[i:i + batch_size]]) for i in range(0, len(Xtr), batch_size)
                      if len(perm[i:i + batch_size]) > 1)
            train_losses.append(train_one_epoch(model, loader, criterion, optimizer))
            scheduler.step()
            in_swa = use_swa and epoch >= swa_start
            if in_swa:
                model.swa_update()
                swa_scheduler.step()
                early_stop_active = True
            val_loss, y_val_pred = evaluate(model, [(Xva_d, yva_t)], criterion, swa=in_swa)
            val_losses.append(val_loss)
            _, _, rmse_val, _, krcc_val = compute_correlation_metrics(yva, y_val_pred)
            metric = rmse_val if select_criteria == 'byrmse' else krcc_val
            if (select_criteria == 'byrmse' and metric < best_metric) or (select_criteria == 'bykrcc' and metric > best_metric):
                best_metric, best_sd, fold_best = metric, model.state_dict(swa=in_swa), model
            if early_stop_active:
                if val_loss < best_val_loss:
                    best_val_loss, epochs_no_improve = val_loss, 0
                    best_sd = model.state_dict(swa=False)                                    # ref :437-441 keeps the plain model here
                else:
                    epochs_no_improve += 1
                    if epochs_no_improve >= patience:
                        break
        all_train_losses.append(train_losses)
        all_val_losses.append(val_losses)
        model.close()
    return best_sd, all_train_losses, all_val_losses
