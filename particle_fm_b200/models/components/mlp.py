"""Conditional MLP vector fields of the jet-feature flow -- mirror of particle_fm/models/components/mlp.py.

``MLP`` (:5-21) and ``small_cond_MLP_model`` (:24-68) keep the reference's constructors, module structure and
state_dict keys (``mlp1.0.weight`` ... ``mlp4.4.bias``).  Without autograd the forward is ONE fused CUDA kernel
(``pfm_mlp_forward``; sampling integrates all steps in one launch, ``pfm_mlp_sample``).  With autograd enabled
(training of this 0.4 M-parameter net) the same parameters are evaluated through torch ops on the device -- the
training of the jet-feature model is not part of the fused path (SURVEY 8f3 names sampling and the step-1 -> step-2
chaining).  There is no CPU path for sampling."""
from __future__ import annotations

from typing import List

import torch
from torch import nn

from ...engine import MlpFlowEngine


class MLP(nn.Sequential):
    def __init__(self, in_features: int, out_features: int, hidden_features: List[int] = [64, 64], activation: str = "ELU"):
        layers = []
        for a, b in zip([in_features] + list(hidden_features), list(hidden_features) + [out_features]):
            layers.extend([nn.Linear(a, b), getattr(nn, activation)()])
        super().__init__(*layers[:-1])


class _CondMLPBase(nn.Module):
    """Blocks of ``MLP`` whose inputs are ``torch.cat([t, x, cond])``; subclasses define ``self.blocks()``."""

    activation: str
    dim_t: int
    dim_cond: int
    in_features: int

    def blocks(self) -> List[MLP]:
        raise NotImplementedError

    def linears(self) -> List[nn.Linear]:
        return [m for blk in self.blocks() for m in blk if isinstance(m, nn.Linear)]

    def _program(self):
        widths, concat, act = [], [], []
        for blk in self.blocks():
            lins = [m for m in blk if isinstance(m, nn.Linear)]
            for i, lin in enumerate(lins):
                widths.append(lin.out_features)
                concat.append(1 if i == 0 else 0)
                act.append(1 if i + 1 < len(lins) else 0)
        return widths, concat, act

    def _weights_key(self):
        from ...engine import weights_generation
        return tuple((p._version, p.data_ptr()) for p in self.parameters()) + (weights_generation(),)

    def engine(self, device=None, force_sync: bool = False) -> MlpFlowEngine:
        p0 = next(self.parameters())
        device = torch.device(device) if device is not None else p0.device
        if device.type != "cuda":
            raise RuntimeError(f"the jet-feature flow's parameters are on {device}: the B200 path needs a CUDA device "
                               "(no CPU fallback; use oracle/ for CPU reference numbers)")
        engines = self.__dict__.setdefault("_engines", {})
        idx = device.index if device.index is not None else torch.cuda.current_device()
        eng = engines.get(idx)
        if eng is None:
            widths, concat, act = self._program()
            eng = MlpFlowEngine(self.in_features, self.dim_t, self.dim_cond, widths, concat, act, self.activation,
                                torch.device("cuda", idx))
            engines[idx] = eng
        key = self._weights_key()
        if force_sync or eng.weights_key != key:
            lins = self.linears()
            eng.set_weights([l.weight for l in lins], [l.bias for l in lins], key=key)
        return eng

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_engines", None)
        return st

    def forward(self, t, x, cond):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_grad and x.is_cuda:
            return self.engine(x.device).forward(t, x, cond)
        if not x.is_cuda:
            raise RuntimeError("the jet-feature flow runs on CUDA devices only (no CPU fallback)")
        for blk in self.blocks():                    # training: torch ops on the device (mlp.py:58-68)
            x = blk(torch.cat([t, x, cond], dim=-1))
        return x


class small_cond_MLP_model(_CondMLPBase):
    def __init__(self, in_features: int, out_features: int, activation: str = "ELU", dim_t: int = 6, dim_cond: int = 1):
        super().__init__()
        if in_features != out_features:
            raise NotImplementedError("the CUDA path integrates a vector field: in_features must equal out_features")
        self.in_features, self.activation, self.dim_t, self.dim_cond = in_features, activation, dim_t, dim_cond
        self.mlp1 = MLP(in_features + dim_t + dim_cond, out_features=64, hidden_features=[64, 64], activation=activation)
        self.mlp2 = MLP(64 + dim_t + dim_cond, out_features=256, hidden_features=[256, 256], activation=activation)
        self.mlp3 = MLP(256 + dim_t + dim_cond, out_features=256, hidden_features=[256, 256], activation=activation)
        self.mlp4 = MLP(256 + dim_t + dim_cond, out_features=out_features, hidden_features=[64, 64], activation=activation)

    def blocks(self):
        return [self.mlp1, self.mlp2, self.mlp3, self.mlp4]


class very_small_cond_MLP_model(_CondMLPBase):
    def __init__(self, in_features: int, out_features: int, activation: str = "ELU", dim_t: int = 6, dim_cond: int = 1):
        super().__init__()
        if in_features != out_features:
            raise NotImplementedError("the CUDA path integrates a vector field: in_features must equal out_features")
        self.in_features, self.activation, self.dim_t, self.dim_cond = in_features, activation, dim_t, dim_cond
        self.mlp1 = MLP(in_features + dim_t + dim_cond, out_features=out_features, hidden_features=[64, 64],
                        activation=activation)

    def blocks(self):
        return [self.mlp1]
