"""Oracle (test infrastructure): EPiC vector-field network, CPU fp32, functional form.

Restates, without nn.Module machinery, what the reference computes in
  particle_fm/models/components/epic.py        EPiC_layer.forward :85-203, EPiC_encoder.forward :304-391
  particle_fm/models/components/time_emb.py    cosine_encoding :49-96
  particle_fm/models/flow_matching_module.py   CNF.forward :191-204, CNF.time_embedding :206-233
Weights come as a flat ``state_dict``-style mapping with the reference's key names
(``fc_l1.weight_g``, ``nn_list.0.fc_global1.weight_v`` ...).  The concatenations are kept
literally (one F.linear on the concatenated input) so the arithmetic is the reference's.
Pinned bit-for-bit against the reference modules by oracle/make_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Mapping, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class EpicCfg:
    """Dimensions of one EPiC_encoder (ctor args of epic.py:226-243, resolved)."""

    feats: int = 3            # output features  (``feats``)
    input_dim: int = 3        # per-particle input width (features [+ 2*frequencies if add_time_to_input])
    hid: int = 128            # ``hid_d``
    latent: int = 10
    layers: int = 6           # ``equiv_layers``
    t_dim: int = 32           # 2 * frequencies
    t_local_cat: bool = True
    t_global_cat: bool = True
    global_cond_dim: int = 0
    local_cond_dim: int = 0
    sum_scale: float = 1e-2
    neg_slope: float = 0.01   # F.leaky_relu default, epic.py:180

    def as_dict(self):
        return asdict(self)


LINEAR_NAMES_STEM = ("fc_l1", "fc_l2", "fc_g1", "fc_g2")
LINEAR_NAMES_LAYER = ("fc_global1", "fc_global2", "fc_local1", "fc_local2")


def linear_names(cfg: EpicCfg):
    """All weight-normed linears of the encoder in state_dict order (epic.py:262-300)."""
    names = list(LINEAR_NAMES_STEM)
    for i in range(cfg.layers):
        names += [f"nn_list.{i}.{n}" for n in LINEAR_NAMES_LAYER]
    names.append("fc_l3")
    return names


def linear_shapes(cfg: EpicCfg):
    """(out, in) of every linear, following the constructors epic.py:66-81 and :262-300."""
    tl = cfg.t_dim if cfg.t_local_cat else 0
    tg = cfg.t_dim if cfg.t_global_cat else 0
    cl, cg, H, Z = cfg.local_cond_dim, cfg.global_cond_dim, cfg.hid, cfg.latent
    shp = {
        "fc_l1": (H, cfg.input_dim + tl + cl),
        "fc_l2": (H, H + tl + cl),
        "fc_g1": (H, 2 * H + tg + cg),
        "fc_g2": (Z, H + tg + cg),
        "fc_l3": (cfg.feats, H + tl + cl),
    }
    for i in range(cfg.layers):
        shp[f"nn_list.{i}.fc_global1"] = (H, 2 * H + Z + tg + cg)
        shp[f"nn_list.{i}.fc_global2"] = (Z, H + tg + cg)
        shp[f"nn_list.{i}.fc_local1"] = (H, H + Z + tl + cl)
        shp[f"nn_list.{i}.fc_local2"] = (H, H + tl + cl)
    return shp


def folded_weight(sd: Mapping[str, Tensor], name: str) -> Tensor:
    """Effective weight of a linear.  Old-style ``nn.utils.weight_norm`` (dim=0):
    W = v * (g / ||v||_2 per output row)  (epic.py:66-81 wrap every linear; torch._weight_norm).
    A plain ``.weight`` key (wrapper_func != "weight_norm" falls back to identity, epic.py:66) is used as is."""
    if f"{name}.weight" in sd:
        return sd[f"{name}.weight"]
    v, g = sd[f"{name}.weight_v"], sd[f"{name}.weight_g"]
    return torch._weight_norm(v, g, 0)


def _lin(sd, name, inp):
    return F.linear(inp, folded_weight(sd, name), sd[f"{name}.bias"])


def cosine_time_code(t: Tensor, outp_dim: int = 32) -> Tensor:
    """time_emb.py:49-96 with the CNF's settings (min 0, max 1, exponential frequencies,
    flow_matching_module.py:183-188): cos((t + 0) * exp(arange(D)) * pi / (1 + 0))."""
    if t.shape[-1] != 1 or t.dim() == 1:          # time_emb.py:79-80
        t = t.unsqueeze(-1)
    freqs = torch.arange(outp_dim, device=t.device).exp()      # int64 -> fp32, time_emb.py:90
    return torch.cos((t + 0.0) * freqs * math.pi / (1.0 + 0.0))  # time_emb.py:96


def sincos_time_code(t: Tensor, frequencies: int) -> Tensor:
    """flow_matching_module.py:172 (buffer 2**arange(f) * pi) and :208-210."""
    fr = 2 ** torch.arange(frequencies) * torch.pi
    a = fr.to(t.device) * t[..., None]
    return torch.cat((a.cos(), a.sin()), dim=-1)


def time_embedding(t: Tensor, x: Tensor, t_emb: str, frequencies: int) -> Tensor:
    """CNF.time_embedding, flow_matching_module.py:206-233 -> (B, N, 2*frequencies) (expanded view)."""
    if t_emb == "sincos":
        code = sincos_time_code(t, frequencies)
    elif t_emb == "cosine":
        if t.dim() == 0:                           # sampling passes a 0-dim t, :225-226
            t = t.unsqueeze(0)
        code = cosine_time_code(t, 2 * frequencies)
    else:
        raise NotImplementedError(t_emb)
    return code.expand(*x.shape[:-1], -1)


def epic_forward(sd: Mapping[str, Tensor], cfg: EpicCfg, t_code: Optional[Tensor], x: Tensor,
                 cond: Optional[Tensor] = None, mask: Optional[Tensor] = None) -> Tensor:
    """EPiC_encoder.forward (epic.py:304-391).  t_code: (B,N,T); x: (B,N,input_dim);
    cond: (B,C) or None; mask: (B,N,1) (any dtype) or None.  Returns (B,N,feats)."""
    act = lambda z: F.leaky_relu(z, cfg.neg_slope)
    B, N = x.shape[0], x.shape[1]
    empty_l = x.new_zeros(B, N, 0)
    empty_g = x.new_zeros(B, 0)
    if mask is None:                                            # epic.py:331-332
        mask = torch.ones_like(x[:, :, 0]).unsqueeze(-1)
    tl = t_code if cfg.t_local_cat else empty_l                 # :335-338
    tg = t_code[:, 0, :] if cfg.t_global_cat else empty_g       # :340-344
    cg = cond if cfg.global_cond_dim > 0 else empty_g           # :347-350
    cl = cond.unsqueeze(-2).expand(B, N, cond.shape[-1]) if cfg.local_cond_dim > 0 else empty_l  # :353-357

    h = act(_lin(sd, "fc_l1", torch.cat((tl, x, cl), -1)))      # :360-362
    h = act(_lin(sd, "fc_l2", torch.cat((tl, h, cl), -1)) + h)  # :364-366
    s = (h * mask).sum(1)                                       # :369
    mean = s / mask.sum(1)                                      # :370
    s = s * cfg.sum_scale                                       # :371
    g = torch.cat((s, mean), -1)                                # :373  (sum first, then mean)
    g = act(_lin(sd, "fc_g1", torch.cat((tg, g, cg), -1)))      # :375-377
    g = act(_lin(sd, "fc_g2", torch.cat((tg, g, cg), -1)))      # :378-380
    for i in range(cfg.layers):                                 # :382-385 -> EPiC_layer.forward
        p = f"nn_list.{i}."
        s = (h * mask).sum(-2)                                  # :160
        mean = s / mask.sum(-2)                                 # :161
        s = s * cfg.sum_scale                                   # :162
        pooled = torch.cat((mean, s, g), -1)                    # :164-171 (mean, sum, global)
        g1 = act(_lin(sd, p + "fc_global1", torch.cat((tg, pooled, cg), -1)))       # :180-182
        g = act(_lin(sd, p + "fc_global2", torch.cat((tg, g1, cg), -1)) + g)        # :184-186
        gb = g.unsqueeze(-2).expand(B, N, g.shape[-1])                               # :189
        u = act(_lin(sd, p + "fc_local1", torch.cat((tl, h, gb, cl), -1)))          # :190-196
        h = act(_lin(sd, p + "fc_local2", torch.cat((tl, u, cl), -1)) + h)          # :198-200
    out = act(_lin(sd, "fc_l3", torch.cat((tl, h, cl), -1)))    # :387-389
    return out * mask                                           # :391


def cnf_forward(sd: Mapping[str, Tensor], cfg: EpicCfg, t: Tensor, x: Tensor, cond=None, mask=None,
                t_emb: str = "cosine", frequencies: int = 16, add_time_to_input: bool = False) -> Tensor:
    """CNF.forward (flow_matching_module.py:191-204): embed t, optional cat((t, x)), run the net."""
    code = time_embedding(t, x, t_emb, frequencies)
    if add_time_to_input:
        x = torch.cat((code, x), dim=-1)                        # :199-200 (time first)
    return epic_forward(sd, cfg, code, x, cond, mask)


# ----------------------------------------------------------------------------------------------
# deterministic, platform-independent synthetic weights / inputs (shared by golden generator + tests)
# ----------------------------------------------------------------------------------------------
def synth_state_dict(cfg: EpicCfg, seed: int, weight_norm: bool = True, g_jitter: float = 0.25):
    """Weights drawn with numpy's legacy RandomState (stream-stable across platforms): v, bias ~
    U(+-1/sqrt(in)) like nn.Linear's default init; g = ||v|| * (1 + jitter*U(-1,1)) so that the
    weight-norm fold is actually exercised (at reference init g == ||v||)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    sd = {}
    shapes = linear_shapes(cfg)
    for name in linear_names(cfg):
        o, i = shapes[name]
        k = 1.0 / math.sqrt(i)
        v = torch.from_numpy(rs.uniform(-k, k, size=(o, i)).astype("float32"))
        b = torch.from_numpy(rs.uniform(-k, k, size=(o,)).astype("float32"))
        if weight_norm:
            jit = torch.from_numpy(rs.uniform(-1, 1, size=(o, 1)).astype("float32"))
            sd[f"{name}.weight_g"] = v.norm(dim=1, keepdim=True) * (1 + g_jitter * jit)
            sd[f"{name}.weight_v"] = v
        else:
            sd[f"{name}.weight"] = v
        sd[f"{name}.bias"] = b
    return sd


def synth_cloud(B: int, N: int, Fdim: int, seed: int, cond_dim: int = 0, all_real: bool = False,
                ragged: bool = False):
    """(x, mask, cond): x ~ N(0,1)*mask; prefix masks with n_real ~ randint(max(1,N//10), N+1)
    (SURVEY 8d); ``ragged`` scatters the real particles instead of a prefix."""
    import numpy as np
    rs = np.random.RandomState(seed)
    n_real = np.full(B, N) if all_real else rs.randint(max(1, N // 10), N + 1, size=B)
    mask = np.zeros((B, N, 1), dtype="float32")
    for b in range(B):
        if ragged:
            mask[b, rs.permutation(N)[: n_real[b]], 0] = 1
        else:
            mask[b, : n_real[b], 0] = 1
    x = rs.standard_normal((B, N, Fdim)).astype("float32") * mask
    cond = rs.standard_normal((B, cond_dim)).astype("float32") if cond_dim else None
    return (torch.from_numpy(x), torch.from_numpy(mask),
            None if cond is None else torch.from_numpy(cond))
