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
