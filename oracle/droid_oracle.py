"""Oracle (test infrastructure): the PC-Droid-style set transformers, CPU fp32, functional form.

Restates what the reference computes in particle_fm/models/components/droid_transformer.py for the two
networks the model YAMLs configure (configs/model/fm_droid_transformer.yaml, fm_droid_crossattention.yaml):
  FullTransformerEncoder.forward      :529-548   (TransformerEncoder :433-437, TransformerEncoderLayer :331-344)
  FullCrossAttentionEncoder.forward   :696-711   (CrossAttentionEncoder :602-619, TransformerCrossAttentionLayer :386-397)
  MultiHeadedAttentionBlock.forward   :211-284   (merge_masks :16-52, torch SDPA with a key-padding mask)
  DenseNetwork.forward / MLPBlock     :958-981, :794-812   with the YAML settings: one hidden block,
      Linear(in [+ctxt]) -> LeakyReLU(0.1) -> LayerNorm -> Linear(out)          (get_act "lrlu" :1022-1023)
and CNF.forward for these models (flow_matching_module.py:148-161, :191-204): the time code is concatenated to
the per-particle input (add_time_to_input) AND, through t[:, 0], to the context vector.
Weights come as a flat mapping with the reference's state_dict key names.  Pinned against the reference's own
modules by oracle/make_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Mapping, Optional, Tuple

import torch
import torch.nn.functional as F

from . import epic_oracle as eo

Tensor = torch.Tensor
LRLU = 0.1          # nn.LeakyReLU(0.1), droid_transformer.py:1022-1023
LN_EPS = 1e-5       # nn.LayerNorm default


@dataclass
class DroidCfg:
    kind: str = "full"          # "full" = FullTransformerEncoder, "cross" = FullCrossAttentionEncoder
    feats: int = 3              # outp_dim
    t_dim: int = 32             # 2 * frequencies
    cond_dim: int = 0           # global_cond_dim
    add_time_to_input: bool = True
    model_dim: int = 256
    num_layers: int = 3
    num_heads: int = 16
    ctxt_out: int = 64          # ctxt_embd_config.outp_dim
    embd_hddn: int = 512        # node / ctxt / outp embedders: 2 * model_dim  (:481-490, :657-666)
    dense_hddn: int = 512       # per-layer FFN: 2 * model_dim, or cae_config.dense_config.hddn_dim
    num_tokens: int = 4         # CrossAttentionEncoder default :571

    @property
    def inpt_dim(self) -> int:
        return self.feats + (self.t_dim if self.add_time_to_input else 0)

    @property
    def ctxt_dim(self) -> int:
        return self.cond_dim + self.t_dim      # flow_matching_module.py:153,160

    def as_dict(self):
        return asdict(self)


def yaml_cfg(kind: str, feats: int = 3, cond_dim: int = 0) -> DroidCfg:
    """The shipped model YAMLs."""
    if kind == "full":
        return DroidCfg("full", feats, 32, cond_dim, True, 256, 3, 16, 64, 512, 512)
    return DroidCfg("cross", feats, 32, cond_dim, True, 128, 8, 16, 64, 256, 256)


# ----------------------------------------------------------------------------------------------
# parameter inventory (reference state_dict keys of ``CNF.net``)
# ----------------------------------------------------------------------------------------------
def _dense(prefix: str, inpt: int, ctxt: int, hddn: int, outp: int) -> List[Tuple[str, Tuple[int, ...]]]:
    return [(prefix + "input_block.block.0.weight", (hddn, inpt + ctxt)), (prefix + "input_block.block.0.bias", (hddn,)),
            (prefix + "input_block.block.2.weight", (hddn,)), (prefix + "input_block.block.2.bias", (hddn,)),
            (prefix + "output_block.block.0.weight", (outp, hddn)), (prefix + "output_block.block.0.bias", (outp,))]


def _ln(prefix: str, d: int):
    return [(prefix + "weight", (d,)), (prefix + "bias", (d,))]


def _lin(prefix: str, o: int, i: int):
    return [(prefix + "weight", (o, i)), (prefix + "bias", (o,))]


def param_shapes(cfg: DroidCfg) -> List[Tuple[str, Tuple[int, ...]]]:
    D, C = cfg.model_dim, cfg.ctxt_out
    out = _dense("ctxt_emdb.", cfg.ctxt_dim, 0, cfg.embd_hddn, C)
    if cfg.kind == "full":
        for i in range(cfg.num_layers):
            p = f"te.layers.{i}."
            out += _lin(p + "self_attn.all_linear.", 3 * D, D) + _ln(p + "self_attn.layer_norm.", D)
            out += _lin(p + "self_attn.out_linear.", D, D) + _dense(p + "dense.", D, C, cfg.dense_hddn, D)
            out += _ln(p + "norm1.", D) + _ln(p + "norm2.", D)
        out += _ln("te.final_norm.", D)
    else:
        out += [("cae.global_tokens", (1, cfg.num_tokens, D))]
        for grp in ("from_layers", "to_layers"):
            for i in range(cfg.num_layers):
                p = f"cae.{grp}.{i}."
                for n in ("q_linear", "k_linear", "v_linear"):
                    out += _lin(p + f"cross_attn.{n}.", D, D)
                out += _ln(p + "cross_attn.layer_norm.", D) + _lin(p + "cross_attn.out_linear.", D, D)
                out += _dense(p + "dense.", D, C, cfg.dense_hddn, D)
                out += _ln(p + "norm0.", D) + _ln(p + "norm1.", D) + _ln(p + "norm2.", D)
    out += _dense("node_embd.", cfg.inpt_dim, C, cfg.embd_hddn, D)
    out += _dense("outp_embd.", D, C, cfg.embd_hddn, cfg.feats)
    return out


def synth_state_dict(cfg: DroidCfg, seed: int) -> Dict[str, Tensor]:
    """Deterministic weights (numpy legacy RandomState).  Linear weights/biases ~ U(+-1/sqrt(in)) like nn.Linear's
    default; LayerNorm gains 1 + 0.2 U(-1,1), shifts 0.1 U(-1,1); the layers the reference zero-initialises
    (out_linear, dense / outp output blocks: init_zeros / output_init_zeros) get random values too, otherwise a
    fresh network outputs exactly 0."""
    import numpy as np
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shp in param_shapes(cfg):
        if name == "cae.global_tokens":
            v = rs.standard_normal(shp)
        elif len(shp) == 2:
            k = 1.0 / math.sqrt(shp[1])
            v = rs.uniform(-k, k, size=shp)
        elif ".block.2." in name or "norm" in name:
            v = 1.0 + 0.2 * rs.uniform(-1, 1, size=shp) if name.endswith("weight") else 0.1 * rs.uniform(-1, 1, size=shp)
        else:
            fan_in = dict(param_shapes(cfg))[name[:-4] + "weight"][1]
            k = 1.0 / math.sqrt(fan_in)
            v = rs.uniform(-k, k, size=shp)
        sd[name] = torch.from_numpy(np.asarray(v, dtype="float32"))
    return sd


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def _ln_apply(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], LN_EPS)


def dense_network(sd: Mapping[str, Tensor], p: str, x: Tensor, ctxt: Optional[Tensor]) -> Tensor:
    """DenseNetwork.forward with one hidden block (:958-981): context broadcast over the sequence and concatenated to
    the INPUT of the first linear only (MLPBlock.forward :803, ctxt_in_inpt)."""
    if ctxt is not None:
        while ctxt.dim() < x.dim():
            ctxt = ctxt.unsqueeze(1)
        ctxt = ctxt.expand(*x.shape[:-1], -1)
        x = torch.cat([x, ctxt], dim=-1)
    h = F.linear(x, sd[p + "input_block.block.0.weight"], sd[p + "input_block.block.0.bias"])
    h = F.leaky_relu(h, LRLU)
    h = _ln_apply(sd, p + "input_block.block.2.", h)
    return F.linear(h, sd[p + "output_block.block.0.weight"], sd[p + "output_block.block.0.bias"])


def _heads(x: Tensor, H: int) -> Tensor:
    B, S, D = x.shape
    return x.view(B, S, H, D // H).transpose(1, 2)


def mha(sd, p: str, cfg: DroidCfg, q: Tensor, k: Optional[Tensor], kv_mask: Optional[Tensor], self_attn: bool) -> Tensor:
    """MultiHeadedAttentionBlock.forward (:211-284): v defaults to k; only keys are masked (merge_masks :33-42);
    LayerNorm BEFORE the output projection (do_layer_norm :280-284)."""
    B, L, D = q.shape
    if self_attn:
        qo, ko, vo = F.linear(q, sd[p + "all_linear.weight"], sd[p + "all_linear.bias"]).chunk(3, -1)
    else:
        qo = F.linear(q, sd[p + "q_linear.weight"], sd[p + "q_linear.bias"])
        ko = F.linear(k, sd[p + "k_linear.weight"], sd[p + "k_linear.bias"])
        vo = F.linear(k, sd[p + "v_linear.weight"], sd[p + "v_linear.bias"])
    merged = None
    if kv_mask is not None:
        merged = kv_mask.unsqueeze(-2).expand(-1, L, -1).unsqueeze(1)
    a = F.scaled_dot_product_attention(_heads(qo, cfg.num_heads), _heads(ko, cfg.num_heads), _heads(vo, cfg.num_heads),
                                       attn_mask=merged, dropout_p=0.0)
    a = a.transpose(1, 2).contiguous().view(B, -1, D)
    a = _ln_apply(sd, p + "layer_norm.", a)
    return F.linear(a, sd[p + "out_linear.weight"], sd[p + "out_linear.bias"])


def droid_forward(sd: Mapping[str, Tensor], cfg: DroidCfg, t_code: Tensor, x: Tensor, cond: Optional[Tensor],
                  mask: Tensor) -> Tensor:
    """FullTransformerEncoder.forward / FullCrossAttentionEncoder.forward.  t_code (B,N,T); x (B,N,inpt_dim) already
    holding the time code if add_time_to_input; cond (B,Cg) or None (= empty); mask (B,N,1).  The output is NOT masked."""
    B = x.shape[0]
    m = mask.squeeze(-1).bool()
    cond = x.new_zeros(B, 0) if cond is None else cond
    ctxt = torch.cat([t_code[:, 0], cond], dim=-1)                       # :541 / :706
    ctxt = dense_network(sd, "ctxt_emdb.", ctxt, None)
    h = dense_network(sd, "node_embd.", x, ctxt)
    if cfg.kind == "full":
        for i in range(cfg.num_layers):
            p = f"te.layers.{i}."
            h = h + mha(sd, p + "self_attn.", cfg, _ln_apply(sd, p + "norm1.", h), None, m, True)          # :339-342
            h = h + dense_network(sd, p + "dense.", _ln_apply(sd, p + "norm2.", h), ctxt)                    # :343
        h = _ln_apply(sd, "te.final_norm.", h)                                                               # :437
    else:
        tok = sd["cae.global_tokens"].expand(B, cfg.num_tokens, cfg.model_dim)                               # :611
        for i in range(cfg.num_layers):
            for grp in ("from_layers", "to_layers"):
                p = f"cae.{grp}.{i}."
                qs, kv, km = (tok, h, m) if grp == "from_layers" else (h, tok, None)                       # :615-617
                qs = qs + mha(sd, p + "cross_attn.", cfg, _ln_apply(sd, p + "norm1.", qs), _ln_apply(sd, p + "norm0.", kv),
                              km, False)                                                                      # :394
                qs = qs + dense_network(sd, p + "dense.", _ln_apply(sd, p + "norm2.", qs), ctxt)            # :395
                if grp == "from_layers":
                    tok = qs
                else:
                    h = qs
    return dense_network(sd, "outp_embd.", h, ctxt)


def cnf_forward(sd: Mapping[str, Tensor], cfg: DroidCfg, t: Tensor, x: Tensor, cond=None, mask=None,
                t_emb: str = "cosine", frequencies: int = 16) -> Tensor:
    """CNF.forward for the droid models (flow_matching_module.py:191-204)."""
    code = eo.time_embedding(t, x, t_emb, frequencies)
    xin = torch.cat((code, x), dim=-1) if cfg.add_time_to_input else x
    return droid_forward(sd, cfg, code, xin, cond, mask)
