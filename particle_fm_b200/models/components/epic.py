"""EPiC vector-field network -- host-side mirror of particle_fm/models/components/epic.py.

Same constructor arguments, parameter names and state_dict layout as the reference's
``EPiC_layer`` (epic.py:17-203) and ``EPiC_encoder`` (epic.py:206-391):
``fc_l1.{bias,weight_g,weight_v}``, ``nn_list.{i}.fc_{global1,global2,local1,local2}.*`` ...,
so checkpoints and the EMA callback's load_state_dict swaps keep working.  The modules hold
parameters only; ``EPiC_encoder.forward`` runs the whole network (stem, all EPiC layers, head) in
one fused CUDA kernel of libpfm_b200.so.  There is no PyTorch fallback.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from ...engine import EpicDims, EpicEngine

Tensor = torch.Tensor


class _WNLinear(nn.Module):
    """Parameters of ``nn.utils.weight_norm(nn.Linear(in, out))`` (old hook API, dim=0) with the
    reference's names and registration order (bias, weight_g, weight_v) and the same default init
    (same RNG consumption as constructing the nn.Linear, then g = ||v|| per output row)."""

    def __init__(self, in_features: int, out_features: int, weight_norm: bool = True):
        super().__init__()
        self.in_features, self.out_features, self.weight_norm = in_features, out_features, weight_norm
        lin = nn.Linear(in_features, out_features)
        if weight_norm:
            self.bias = nn.Parameter(lin.bias.detach().clone())
            self.weight_g = nn.Parameter(torch.norm_except_dim(lin.weight.detach(), 2, 0).clone())
            self.weight_v = nn.Parameter(lin.weight.detach().clone())
        else:
            self.weight = nn.Parameter(lin.weight.detach().clone())
            self.bias = nn.Parameter(lin.bias.detach().clone())

    def folded(self) -> Tuple[Tensor, Tensor]:
        """(W, b) with W = v * (g / ||v||) -- differentiable, so autograd maps the kernel's dW back
        onto weight_g / weight_v."""
        if self.weight_norm:
            return torch._weight_norm(self.weight_v, self.weight_g, 0), self.bias
        return self.weight, self.bias

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, weight_norm={self.weight_norm}"

    def forward(self, *a, **k):
        raise NotImplementedError("evaluated inside the fused EPiC kernel, not as a separate op")


def _wrapper_is_weight_norm(wrapper_func: str) -> bool:
    """The reference does ``getattr(nn.utils, wrapper_func, lambda x: x)`` (epic.py:66): a name that
    nn.utils does not have degrades to plain linears; any other real wrapper is unsupported here."""
    if wrapper_func == "weight_norm":
        return True
    if hasattr(nn.utils, wrapper_func):
        raise NotImplementedError(f"wrapper_func={wrapper_func!r}: only 'weight_norm' (or a name torch.nn.utils does "
                                  "not define, i.e. plain linears) is supported by the CUDA path")
    return False


class EPiC_layer(nn.Module):
    """Parameter container for one EPiC layer (constructor of epic.py:37-83)."""

    def __init__(self, local_in_dim: int = 3, hid_dim: int = 256, latent_dim: int = 16, global_cond_dim: int = 0,
                 local_cond_dim: int = 0, t_local_cat: bool = False, t_global_cat: bool = False,
                 activation: str = "leaky_relu", wrapper_func: str = "weight_norm", frequencies: int = 6,
                 num_points: int = 30, dropout: float = 0.0, sum_scale: float = 1e-2):
        super().__init__()
        if activation != "leaky_relu":
            raise NotImplementedError(f"activation={activation!r}: the CUDA path implements leaky_relu (all configs)")
        if dropout != 0.0:
            raise NotImplementedError("dropout > 0 is not supported by the CUDA path (p=0 in every config)")
        self.activation, self.global_cond_dim, self.local_cond_dim = activation, global_cond_dim, local_cond_dim
        self.num_points, self.sum_scale = num_points, sum_scale
        self.t_local_cat, self.t_global_cat = t_local_cat, t_global_cat
        tl = 2 * frequencies if t_local_cat else 0
        tg = 2 * frequencies if t_global_cat else 0
        wn = _wrapper_is_weight_norm(wrapper_func)
        self.fc_global1 = _WNLinear(int(2 * hid_dim) + latent_dim + tg + global_cond_dim, hid_dim, wn)
        self.fc_global2 = _WNLinear(hid_dim + tg + global_cond_dim, latent_dim, wn)
        self.fc_local1 = _WNLinear(local_in_dim + latent_dim + tl + local_cond_dim, hid_dim, wn)
        self.fc_local2 = _WNLinear(hid_dim + tl + local_cond_dim, hid_dim, wn)

    def forward(self, *a, **k):
        raise NotImplementedError("EPiC_layer runs fused inside EPiC_encoder.forward (one kernel for all layers)")


class EPiC_encoder(nn.Module):
    """EPiC encoder used as the flow-matching vector field (epic.py:206-391)."""

    def __init__(self, latent: int = 16, input_dim: int = 3, hid_d: int = 256, feats: int = 128,
                 equiv_layers: int = 8, global_cond_dim: int = 0, local_cond_dim: int = 0,
                 activation: str = "leaky_relu", wrapper_func: str = "weight_norm", frequencies: int = 6,
                 num_points: int = 30, t_local_cat: bool = False, t_global_cat: bool = False, dropout: float = 0.0,
                 sum_scale: float = 1e-2):
        super().__init__()
        if activation != "leaky_relu":
            raise NotImplementedError(f"activation={activation!r}: the CUDA path implements leaky_relu (all configs)")
        if dropout != 0.0:
            raise NotImplementedError("dropout > 0 is not supported by the CUDA path (p=0 in every config)")
        if local_cond_dim not in (0, global_cond_dim) and global_cond_dim != 0:
            raise ValueError("local_cond_dim must be 0 or equal to global_cond_dim: one cond tensor feeds both "
                             "(epic.py:347-357)")
        self.activation, self.latent, self.input_dim, self.hid_d, self.feats = activation, latent, input_dim, hid_d, feats
        self.equiv_layers, self.global_cond_dim, self.local_cond_dim = equiv_layers, global_cond_dim, local_cond_dim
        self.num_points, self.sum_scale = num_points, sum_scale
        self.t_local_cat, self.t_global_cat = t_local_cat, t_global_cat
        self.frequencies = frequencies
        tl = 2 * frequencies if t_local_cat else 0
        tg = 2 * frequencies if t_global_cat else 0
        wn = _wrapper_is_weight_norm(wrapper_func)
        self.fc_l1 = _WNLinear(input_dim + tl + local_cond_dim, hid_d, wn)
        self.fc_l2 = _WNLinear(hid_d + tl + local_cond_dim, hid_d, wn)
        self.fc_g1 = _WNLinear(int(2 * hid_d) + tg + global_cond_dim, hid_d, wn)
        self.fc_g2 = _WNLinear(hid_d + tg + global_cond_dim, latent, wn)
        self.nn_list = nn.ModuleList()
        for _ in range(equiv_layers):
            self.nn_list.append(EPiC_layer(hid_d, hid_d, latent, activation=activation, wrapper_func=wrapper_func,
                                           num_points=num_points, t_global_cat=t_global_cat, t_local_cat=t_local_cat,
                                           global_cond_dim=global_cond_dim, local_cond_dim=local_cond_dim,
                                           frequencies=frequencies, dropout=dropout, sum_scale=sum_scale))
        self.fc_l3 = _WNLinear(hid_d + tl + local_cond_dim, feats, wn)
        # arithmetic of the per-particle contractions: "fp32" (CUDA cores) or "bf16" (tcgen05 tensor cores)
        self.precision = os.environ.get("PFM_PRECISION", "fp32")
        self._engines: Dict[int, EpicEngine] = {}

    # ------------------------------------------------------------------------------------------
    def dims(self) -> EpicDims:
        return EpicDims(feats=self.feats, input_dim=self.input_dim, hid=self.hid_d, latent=self.latent,
                        layers=self.equiv_layers, t_dim=2 * self.frequencies, t_local_cat=self.t_local_cat,
                        t_global_cat=self.t_global_cat, global_cond_dim=self.global_cond_dim,
                        local_cond_dim=self.local_cond_dim, sum_scale=self.sum_scale, neg_slope=0.01)

    def linears(self) -> List[_WNLinear]:
        """State-dict order, the order libpfm_b200 expects."""
        out = [self.fc_l1, self.fc_l2, self.fc_g1, self.fc_g2]
        for layer in self.nn_list:
            out += [layer.fc_global1, layer.fc_global2, layer.fc_local1, layer.fc_local2]
        out.append(self.fc_l3)
        return out

    def _weights_key(self):
        from ...engine import weights_generation
        # every parameter lives in one of linears(); walking their _parameters dicts is ~5x cheaper per training
        # step than Module.parameters() (which recurses through named_modules with a de-duplication set)
        return tuple((p._version, p.data_ptr()) for lin in self.linears() for p in lin._parameters.values()
                     if p is not None) + (weights_generation(),)

    def invalidate_weights(self):
        """Force the next engine() call to re-fold and repack the parameters.  The change detector below keys on
        (``_version``, ``data_ptr``) of every parameter, which in-place updates through ``.data`` (``p.data.mul_()``,
        manual EMA / clipping / re-initialisation) do NOT bump -- call this after such an update.  Sampling entry points
        (``CNF.decode``) re-sync unconditionally, so only per-step forward / training callers need it."""
        for eng in self._engines.values():
            eng.weights_key = None

    def engine(self, device: Optional[torch.device] = None, sync_weights: bool = True,
               force_sync: bool = False) -> EpicEngine:
        """The packed copy of this network on ``device`` (default: where the parameters live), refreshed
        whenever a parameter changed (optimizer step, EMA load_state_dict swap -- ema.py:145-159) as far as
        ``_version`` / ``data_ptr`` show it (see invalidate_weights), or unconditionally with ``force_sync``."""
        p0 = next(self.parameters())
        device = torch.device(device) if device is not None else p0.device
        if device.type != "cuda":
            raise RuntimeError(f"EPiC_encoder parameters are on {device}: the B200 path needs a CUDA device "
                               "(no CPU fallback; use oracle/ for CPU reference numbers)")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        eng = self._engines.get(idx)
        if eng is None:
            eng = EpicEngine(self.dims(), torch.device("cuda", idx), self.precision)
            self._engines[idx] = eng
        if eng.precision != self.precision:
            eng.set_precision(self.precision)
        if sync_weights:
            key = self._weights_key()
            if force_sync or eng.weights_key != key:
                eng.set_params(self.linears(), key=key)      # weight-norm fold + repack inside the library
        return eng

    def _apply(self, fn, *a, **k):      # .to()/.cuda(): parameters move, packed copies are rebuilt lazily
        out = super()._apply(fn, *a, **k)
        for eng in self._engines.values():
            eng.weights_key = None
        return out

    def __getstate__(self):             # engines hold raw CUDA handles: never pickle / deepcopy them
        st = self.__dict__.copy()
        st["_engines"] = {}
        return st

    # ------------------------------------------------------------------------------------------
    def forward(self, t_in: Tensor = None, x_local: Tensor = None, global_cond_in: Tensor = None,
                mask: Tensor = None) -> Tensor:
        """Same contract as epic.py:304-391: t_in (B,N,T), x_local (B,N,input_dim), cond (B,C), mask (B,N,1)."""
        if x_local is None:
            raise ValueError("x_local is None")
        if global_cond_in is None and (self.global_cond_dim > 0 or self.local_cond_dim > 0):
            raise ValueError(f"global_cond_dim is {self.global_cond_dim} and local_cond_dim is {self.local_cond_dim} "
                             "but no global_cond is given")
        if t_in is None and (self.t_local_cat or self.t_global_cat):
            raise ValueError(f"t_local_cat is {self.t_local_cat} and t_global_cat is {self.t_global_cat} but no t is given")
        t_code = None
        if t_in is not None and (self.t_local_cat or self.t_global_cat):
            # the code is constant over the particles of a jet (CNF.time_embedding expands it), and over the
            # batch too when sampling: hand the kernel [1,T] or [B,T]
            t_code = t_in[:, 0, :]
            if t_in.stride(0) == 0 or t_in.shape[0] == 1:
                t_code = t_code[:1]
        needs_grad = torch.is_grad_enabled() and (x_local.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            from ...training import epic_forward_autograd
            return epic_forward_autograd(self, t_code, x_local, global_cond_in, mask)
        return self.engine(x_local.device).forward(t_code, x_local, mask, global_cond_in)
