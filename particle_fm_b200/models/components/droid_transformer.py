"""PC-Droid-style set transformers -- host-side mirror of particle_fm/models/components/droid_transformer.py.

Parameter containers with the reference's module tree and state_dict keys (``ctxt_emdb.input_block.block.0.weight``,
``te.layers.0.self_attn.all_linear.weight``, ``cae.global_tokens`` ...) for the configurations the model YAMLs
use (configs/model/fm_droid_transformer.yaml, fm_droid_crossattention.yaml): dense networks with one hidden block
(``Linear -> LeakyReLU(0.1) -> LayerNorm -> Linear``), pre-norm attention blocks with a LayerNorm before the output
projection, no dropout.  ``FullTransformerEncoder.forward`` / ``FullCrossAttentionEncoder.forward``
(droid_transformer.py:529-548, :696-711) run in libpfm_b200.so (csrc/tf_simt.cu); other options raise.
"""
from __future__ import annotations

import ctypes as C
from copy import deepcopy
from typing import Dict, Mapping, Optional

import torch
import torch.nn as nn

from ... import _lib

Tensor = torch.Tensor


def _only(cfg: Mapping, allowed: Mapping, what: str):
    for k, v in cfg.items():
        if k not in allowed:
            raise NotImplementedError(f"{what}: option {k!r} is not supported by the CUDA path")
        if allowed[k] is not None and v not in allowed[k]:
            raise NotImplementedError(f"{what}: {k}={v!r} is not supported by the CUDA path (supported: {allowed[k]})")


class MLPBlock(nn.Module):
    """``block`` = [Linear, LeakyReLU(0.1), LayerNorm] (hidden) or [Linear] (output)   (droid_transformer.py:714-812)."""

    def __init__(self, inpt_dim: int, outp_dim: int, ctxt_dim: int = 0, hidden: bool = True, init_zeros: bool = False):
        super().__init__()
        self.inpt_dim, self.outp_dim, self.ctxt_dim = inpt_dim, outp_dim, ctxt_dim
        self.block = nn.ModuleList([nn.Linear(inpt_dim + ctxt_dim, outp_dim)])
        if init_zeros:
            self.block[0].weight.data.fill_(0)
            self.block[0].bias.data.fill_(0)
        if hidden:
            self.block.append(nn.LeakyReLU(0.1))
            self.block.append(nn.LayerNorm(outp_dim))


class DenseNetwork(nn.Module):
    """One hidden block + output block (droid_transformer.py:828-981 with the YAML settings)."""

    def __init__(self, inpt_dim: int, outp_dim: int = 0, ctxt_dim: int = 0, hddn_dim: int = 32, **cfg):
        super().__init__()
        _only(cfg, dict(act_h=("lrlu",), nrm=("layer",), output_init_zeros=(True, False), num_blocks=(1,), n_lyr_pbk=(1,),
                        act_o=("none",), do_out=(True,), drp=(0, 0.0), drp_on_output=(False,), nrm_on_output=(False,),
                        do_res=(False,), ctxt_in_inpt=(True,), ctxt_in_hddn=(False,)), "DenseNetwork")
        if cfg.get("act_h", "lrlu") != "lrlu" or cfg.get("nrm", "none") != "layer":
            raise NotImplementedError("DenseNetwork: the CUDA path implements act_h='lrlu' with nrm='layer' (the model YAMLs)")
        if not isinstance(hddn_dim, int):
            raise NotImplementedError("DenseNetwork: a list of hidden widths is not supported by the CUDA path")
        self.inpt_dim, self.ctxt_dim, self.hddn_dim = inpt_dim, ctxt_dim, [hddn_dim]
        self.outp_dim = outp_dim or inpt_dim
        self.input_block = MLPBlock(inpt_dim, hddn_dim, ctxt_dim, hidden=True)
        self.output_block = MLPBlock(hddn_dim, self.outp_dim, 0, hidden=False, init_zeros=cfg.get("output_init_zeros", False))


class MultiHeadedAttentionBlock(nn.Module):
    def __init__(self, model_dim: int, num_heads: int = 1, drp: float = 0, init_zeros: bool = False, do_selfattn: bool = False,
                 do_layer_norm: bool = False, attn_act=None):
        super().__init__()
        if drp or attn_act is not None or not do_layer_norm:
            raise NotImplementedError("MultiHeadedAttentionBlock: the CUDA path implements do_layer_norm=True, drp=0, softmax")
        if model_dim % num_heads:
            raise ValueError("Model dimension must be divisible by number of heads!")
        self.model_dim, self.num_heads, self.head_dim, self.do_selfattn = model_dim, num_heads, model_dim // num_heads, do_selfattn
        if do_selfattn:
            self.all_linear = nn.Linear(model_dim, 3 * model_dim)
        else:
            self.q_linear = nn.Linear(model_dim, model_dim)
            self.k_linear = nn.Linear(model_dim, model_dim)
            self.v_linear = nn.Linear(model_dim, model_dim)
        self.layer_norm = nn.LayerNorm(model_dim)
        self.out_linear = nn.Linear(model_dim, model_dim)
        if init_zeros:
            self.out_linear.weight.data.fill_(0)
            self.out_linear.bias.data.fill_(0)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, model_dim: int, mha_config: Mapping = None, dense_config: Mapping = None, ctxt_dim: int = 0):
        super().__init__()
        self.self_attn = MultiHeadedAttentionBlock(model_dim, do_selfattn=True, **(mha_config or {}))
        self.dense = DenseNetwork(model_dim, outp_dim=model_dim, ctxt_dim=ctxt_dim, **(dense_config or {}))
        self.norm1 = nn.LayerNorm(model_dim)
        self.norm2 = nn.LayerNorm(model_dim)


class TransformerCrossAttentionLayer(nn.Module):
    def __init__(self, model_dim: int, mha_config: Mapping = None, dense_config: Mapping = None, ctxt_dim: int = 0):
        super().__init__()
        self.cross_attn = MultiHeadedAttentionBlock(model_dim, do_selfattn=False, **(mha_config or {}))
        self.dense = DenseNetwork(model_dim, outp_dim=model_dim, ctxt_dim=ctxt_dim, **(dense_config or {}))
        self.norm0 = nn.LayerNorm(model_dim)
        self.norm1 = nn.LayerNorm(model_dim)
        self.norm2 = nn.LayerNorm(model_dim)


class TransformerEncoder(nn.Module):
    def __init__(self, model_dim: int = 64, num_layers: int = 3, mha_config: Mapping = None, dense_config: Mapping = None,
                 ctxt_dim: int = 0):
        super().__init__()
        self.model_dim, self.num_layers = model_dim, num_layers
        self.layers = nn.ModuleList([TransformerEncoderLayer(model_dim, mha_config, dense_config, ctxt_dim)
                                     for _ in range(num_layers)])
        self.final_norm = nn.LayerNorm(model_dim)


class CrossAttentionEncoder(nn.Module):
    def __init__(self, model_dim: int = 64, num_tokens: int = 4, num_layers: int = 5, mha_config: Mapping = None,
                 dense_config: Mapping = None, ctxt_dim: int = 0):
        super().__init__()
        self.model_dim, self.num_layers, self.num_tokens = model_dim, num_layers, num_tokens
        self.global_tokens = nn.Parameter(torch.randn((1, num_tokens, model_dim)))
        self.from_layers = nn.ModuleList([TransformerCrossAttentionLayer(model_dim, mha_config, dense_config, ctxt_dim)
                                          for _ in range(num_layers)])
        self.to_layers = nn.ModuleList([TransformerCrossAttentionLayer(model_dim, mha_config, dense_config, ctxt_dim)
                                        for _ in range(num_layers)])


class _TfEngine:
    """One packed copy of a droid network on one GPU (C ABI: pfm_tf_*)."""

    def __init__(self, cfg: Dict, device: torch.device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.cfg = cfg
        c = _lib.TfCfgC(cfg["kind"], cfg["feats"], cfg["t_dim"], cfg["cond_dim"], int(cfg["add_time_to_input"]), cfg["model_dim"],
                        cfg["num_layers"], cfg["num_heads"], cfg["ctxt_out"], cfg["embd_hddn"], cfg["dense_hddn"],
                        cfg["num_tokens"], 0.1, 1e-5)
        h = C.c_void_p()
        _lib.check(self.lib.pfm_tf_create(C.byref(c), self.index, C.byref(h)), "pfm_tf_create")
        self._h = h
        self.weights_key = None
        self.precision = "fp32"

    def set_precision(self, precision: str):
        code = {"fp32": _lib.PFM_PREC_FP32, "bf16": _lib.PFM_PREC_BF16}.get(precision)
        if code is None:
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
        _lib.check(self.lib.pfm_tf_set_precision(self._h, code), "pfm_tf_set_precision")
        self.precision = precision

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self.lib.pfm_tf_destroy(h)
            except Exception:
                pass
            self._h = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def param_shapes(self):
        n = self.lib.pfm_tf_num_params(self._h)
        r, c = C.c_int32(), C.c_int32()
        out = []
        for i in range(n):
            _lib.check(self.lib.pfm_tf_param_shape(self._h, i, C.byref(r), C.byref(c)), "pfm_tf_param_shape")
            out.append((r.value, c.value))
        return out

    def set_weights(self, params, key=None):
        shapes = self.param_shapes()
        if len(params) != len(shapes):
            raise ValueError(f"expected {len(shapes)} parameter tensors, got {len(params)}")
        ps = []
        for i, (p, (r, c)) in enumerate(zip(params, shapes)):
            t = p.detach().to(device=self.device, dtype=torch.float32).contiguous()
            if t.numel() != r * c:
                raise ValueError(f"parameter {i}: expected {r}x{c} = {r * c} values, got shape {tuple(p.shape)}")
            ps.append(t)
        arr = (C.c_void_p * len(ps))(*[t.data_ptr() for t in ps])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_tf_set_weights(self._h, arr, len(ps), self._stream()), "pfm_tf_set_weights")
        self._keep = ps
        self.weights_key = key

    @staticmethod
    def _f32(t, device):
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    def _cond(self, cond, B):
        if self.cfg["cond_dim"] == 0:
            return None
        if cond is None:
            raise ValueError(f"global_cond_dim={self.cfg['cond_dim']} but cond is None")
        cond = self._f32(cond, self.device).reshape(B, -1)
        if cond.shape[1] != self.cfg["cond_dim"]:
            raise ValueError(f"cond has {cond.shape[1]} columns, expected {self.cfg['cond_dim']}")
        return cond

    def forward(self, t_code: Tensor, x: Tensor, mask: Optional[Tensor], cond: Optional[Tensor]) -> Tensor:
        B, N = int(x.shape[0]), int(x.shape[1])
        x = self._f32(x, self.device)
        mask = None if mask is None else self._f32(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        t_code = self._f32(t_code, self.device).reshape(-1, self.cfg["t_dim"])
        out = torch.empty(B, N, self.cfg["feats"], device=self.device, dtype=torch.float32)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_tf_forward(self._h, p(t_code), int(t_code.shape[0]), p(x), p(mask), p(cond), p(out), B, N,
                                               self._stream()), "pfm_tf_forward")
        return out

    def sample(self, z, mask, cond, t_codes, t_codes_in, dt, solver):
        """Same call shape as EpicEngine.sample (t_codes_in is the same table: the library hoists the input time columns)."""
        B, N = int(z.shape[0]), int(z.shape[1])
        x = z.detach().to(device=self.device, dtype=torch.float32).contiguous().clone()
        mask = None if mask is None else self._f32(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        codes = t_codes if t_codes is not None else t_codes_in
        dt = self._f32(dt, self.device).reshape(-1)
        n_steps = int(dt.numel())
        n_evals = n_steps * (2 if solver == "midpoint" else 1)
        codes = self._f32(codes, self.device).reshape(n_evals, self.cfg["t_dim"])
        code = {"euler": _lib.PFM_SOLVER_EULER, "midpoint": _lib.PFM_SOLVER_MIDPOINT}[solver]
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_tf_sample(self._h, p(x), p(mask), p(cond), p(codes), p(dt), code, n_steps, B, N, self._stream()),
                       "pfm_tf_sample")
        return x

    # ---- training (fp32, dense rows: the reference neither masks the output nor the loss of padded slots) ----
    def grad_size(self) -> int:
        return int(self.lib.pfm_tf_grad_size(self._h))

    def grad_views(self, flat: Tensor):
        """Per-parameter views of the flat gradient, in parameter (= state_dict) order, each [rows, cols] row-major."""
        out, off = [], 0
        for r, c in self.param_shapes():
            out.append(flat[off:off + r * c])
            off += r * c
        return out

    def forward_train(self, t_code: Tensor, x: Tensor, mask: Tensor, cond: Optional[Tensor]) -> Tensor:
        B, N = int(x.shape[0]), int(x.shape[1])
        x = self._f32(x, self.device)
        mask = self._f32(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        t_code = self._f32(t_code, self.device).reshape(-1, self.cfg["t_dim"])
        out = torch.empty(B, N, self.cfg["feats"], device=self.device, dtype=torch.float32)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_tf_forward_train(self._h, p(t_code), int(t_code.shape[0]), p(x), p(mask), p(cond), p(out), B, N,
                                                     self._stream()), "pfm_tf_forward_train")
        self._ticket = getattr(self, "_ticket", 0) + 1
        return out, self._ticket

    def backward(self, ticket: int, gout: Tensor) -> Tensor:
        if ticket != getattr(self, "_ticket", None):
            raise RuntimeError("the droid engine keeps ONE saved forward: backward() was called after another training "
                               "forward on the same network")
        gout = self._f32(gout, self.device)
        flat = torch.empty(self.grad_size(), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_tf_backward(self._h, C.c_void_p(gout.data_ptr()), C.c_void_p(flat.data_ptr()), self._stream()),
                       "pfm_tf_backward")
        return flat

    def loss_fwd_bwd(self, kind: str, x, mask, cond, t, t_code, n0, n1, sigma: float, want_grad: bool = True):
        B, N = int(x.shape[0]), int(x.shape[1])
        x = self._f32(x, self.device)
        mask = self._f32(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        t = self._f32(t, self.device).reshape(B)
        t_code = self._f32(t_code, self.device).reshape(B, self.cfg["t_dim"])
        n0 = self._f32(n0, self.device)
        n1 = None if n1 is None else self._f32(n1, self.device)
        loss = torch.empty(1, device=self.device, dtype=torch.float32)
        flat = torch.empty(self.grad_size(), device=self.device, dtype=torch.float32) if want_grad else None
        code = {"FM-OT": _lib.PFM_LOSS_FM_OT, "CFM": _lib.PFM_LOSS_CFM, "droid": _lib.PFM_LOSS_DROID}[kind]
        p = lambda v: None if v is None else C.c_void_p(v.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_tf_loss_fwd_bwd(self._h, p(x), p(t), p(t_code), p(n0), p(n1), p(mask), p(cond), code, float(sigma),
                                                    p(loss), p(flat), B, N, self._stream()), "pfm_tf_loss_fwd_bwd")
        self._ticket = getattr(self, "_ticket", 0) + 1
        return loss, flat

    def grad_chunks(self):
        """The droid backward finishes the flat gradient as a whole: the data-parallel hook all-reduces it in one piece."""
        return []

    def last_launches(self) -> int:
        return int(self.lib.pfm_tf_last_launches(self._h))


class _DroidFn(torch.autograd.Function):
    """Differentiable forward of a droid net: gradients w.r.t. its parameters (not w.r.t. the input)."""

    @staticmethod
    def forward(ctx, eng, t_code, x, mask, cond, *params):
        out, ticket = eng.forward_train(t_code, x, mask, cond)
        ctx.eng, ctx.ticket = eng, ticket
        ctx.shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, gout):
        flat = ctx.eng.backward(ctx.ticket, gout)
        grads = [g.view(shp) if need else None
                 for g, shp, need in zip(ctx.eng.grad_views(flat), ctx.shapes, ctx.needs_input_grad[5:])]
        return (None, None, None, None, None) + tuple(grads)


class _DroidLossFn(torch.autograd.Function):
    """loss = sum((net(t, y) - u)^2) / sum(mask) with the interpolation, forward and backward in one fused call."""

    @staticmethod
    def forward(ctx, net, eng, kind, sigma, x, mask, cond, t, t_code, n0, n1, *params):
        want = any(ctx.needs_input_grad[11:])
        loss, flat = eng.loss_fwd_bwd(kind, x, mask, cond, t, t_code, n0, n1, sigma, want_grad=want)
        hook = getattr(net, "flat_grad_hook", None)
        if want and hook is not None:
            flat = hook(flat, eng)                # e.g. the data-parallel all-reduce of the flat gradient
        ctx.eng, ctx.flat = eng, flat
        ctx.shapes = [p.shape for p in params]
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        head = (None,) * 11
        if ctx.flat is None:
            return head + (None,) * len(ctx.shapes)
        grads = [(gv.view(shp) * g) if need else None
                 for gv, shp, need in zip(ctx.eng.grad_views(ctx.flat), ctx.shapes, ctx.needs_input_grad[11:])]
        return head + tuple(grads)


def droid_loss_autograd(cnf, kind: str, x: Tensor, mask: Tensor, cond: Optional[Tensor], t: Tensor, n0: Tensor,
                        n1: Optional[Tensor], sigma: float) -> Tensor:
    """Scalar flow-matching loss with an autograd graph to the droid net's parameters (t: (B,) per jet)."""
    net = cnf.net
    if x.device.type != "cuda":
        raise RuntimeError(f"the flow-matching loss got a batch on {x.device}: the B200 path needs a CUDA device "
                           "(no CPU fallback; use oracle/ for CPU reference numbers)")
    eng = net.engine(x.device)
    with torch.no_grad():
        code = cnf.time_code(t.to(x.device))
    return _DroidLossFn.apply(net, eng, kind, sigma, x, mask, cond, t, code, n0, n1, *list(net.parameters()))


class _DroidNet(nn.Module):
    """Shared host logic of the two full encoders: packed-weight cache, forward through the C ABI."""
    kind = -1
    t_local_cat = False          # CNF.decode asks the net whether it takes the time code besides the input columns:
    t_global_cat = True          # the droid nets always do (context = [t, cond], droid_transformer.py:541,706)

    def _finish(self, inpt_dim, outp_dim, ctxt_dim, t_dim):
        self.inpt_dim, self.outp_dim, self.ctxt_dim, self.t_dim = inpt_dim, outp_dim, ctxt_dim, t_dim
        self._engines: Dict[int, _TfEngine] = {}
        import os
        self.precision = os.environ.get("PFM_PRECISION", "fp32")

    def _engine_cfg(self) -> Dict:
        core = self.te if self.kind == 0 else self.cae
        layer0 = core.layers[0] if self.kind == 0 else core.from_layers[0]
        mha = layer0.self_attn if self.kind == 0 else layer0.cross_attn
        add_time = self.inpt_dim > self.outp_dim
        return dict(kind=self.kind, feats=self.outp_dim, t_dim=self.t_dim, cond_dim=self.ctxt_dim - self.t_dim,
                    add_time_to_input=add_time, model_dim=core.model_dim, num_layers=core.num_layers, num_heads=mha.num_heads,
                    ctxt_out=self.ctxt_emdb.outp_dim, embd_hddn=self.node_embd.hddn_dim[0], dense_hddn=layer0.dense.hddn_dim[0],
                    num_tokens=getattr(core, "num_tokens", 4))

    def _weights_key(self):
        from ...engine import weights_generation
        return tuple((p._version, p.data_ptr()) for p in self.parameters()) + (weights_generation(),)

    def invalidate_weights(self):
        """See EPiC_encoder.invalidate_weights: in-place ``.data`` updates are invisible to the change detector."""
        for eng in self._engines.values():
            eng.weights_key = None

    def engine(self, device=None, sync_weights: bool = True, force_sync: bool = False) -> _TfEngine:
        p0 = next(self.parameters())
        device = torch.device(device) if device is not None else p0.device
        if device.type != "cuda":
            raise RuntimeError(f"the droid transformer's parameters are on {device}: the B200 path needs a CUDA device "
                               "(no CPU fallback; use oracle/ for CPU reference numbers)")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        eng = self._engines.get(idx)
        if eng is None:
            if self.node_embd.hddn_dim[0] != self.outp_embd.hddn_dim[0] or self.node_embd.hddn_dim[0] != self.ctxt_emdb.hddn_dim[0]:
                raise NotImplementedError("the CUDA path expects one hidden width for the node / ctxt / outp embedders")
            eng = _TfEngine(self._engine_cfg(), torch.device("cuda", idx))
            self._engines[idx] = eng
        if eng.precision != self.precision:
            eng.set_precision(self.precision)
        if sync_weights:
            key = self._weights_key()
            if force_sync or eng.weights_key != key:
                eng.set_weights(list(self.parameters()), key=key)
        return eng

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        for eng in self._engines.values():
            eng.weights_key = None
        return out

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engines"] = {}
        return st

    def forward(self, t: Tensor, x: Tensor, ctxt: Tensor = None, mask: Tensor = None) -> Tensor:
        """t (B,N,T) time code; x (B,N,inpt_dim) (time code first if add_time_to_input); ctxt (B,Cg) or None; mask (B,N,1)."""
        if mask is None:
            raise ValueError("the droid transformers need a mask (the reference calls mask.squeeze(-1).bool())")
        t_code = t[:, 0, :]
        if t.stride(0) == 0 or t.shape[0] == 1:
            t_code = t_code[:1]
        xin = x[..., x.shape[-1] - self.outp_dim:]                   # drop the concatenated time columns: they are hoisted
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if x.requires_grad:
                raise NotImplementedError("gradients w.r.t. the input of the droid transformers are not provided by the B200 "
                                          "path (parameter gradients are)")
            # training semantics of the reference: every slot is evaluated (padded slots are not masked on output)
            return _DroidFn.apply(self.engine(x.device), t_code, xin, mask, ctxt, *list(self.parameters()))
        return self.engine(x.device).forward(t_code, xin, mask, ctxt)


class FullTransformerEncoder(_DroidNet):
    """droid_transformer.py:440-548."""
    kind = 0

    def __init__(self, inpt_dim: int, outp_dim: int, edge_dim: int = 0, ctxt_dim: int = 0, te_config: Mapping = None,
                 node_embd_config: Mapping = None, outp_embd_config: Mapping = None, edge_embd_config: Mapping = None,
                 ctxt_embd_config: Mapping = None):
        super().__init__()
        if edge_dim:
            raise NotImplementedError("edge features (attn_bias) are not supported by the CUDA path")
        if not ctxt_dim:
            raise NotImplementedError("the CUDA path expects the time code in the context (ctxt_dim = cond + 2*frequencies)")
        te_config, node_embd_config = deepcopy(dict(te_config or {})), deepcopy(dict(node_embd_config or {}))
        outp_embd_config, ctxt_embd_config = deepcopy(dict(outp_embd_config or {})), deepcopy(dict(ctxt_embd_config or {}))
        te_config["dense_config"] = dict(te_config.get("dense_config") or {})
        if "model_dim" in te_config:                               # dense networks double the width by default (:481-490)
            md = te_config["model_dim"]
            for cfg in (node_embd_config, ctxt_embd_config, outp_embd_config, te_config["dense_config"]):
                cfg.setdefault("hddn_dim", 2 * md)
        self.ctxt_emdb = DenseNetwork(inpt_dim=ctxt_dim, **ctxt_embd_config)
        self.ctxt_out = self.ctxt_emdb.outp_dim
        self.te = TransformerEncoder(**te_config, ctxt_dim=self.ctxt_out)
        self.model_dim = self.te.model_dim
        self.node_embd = DenseNetwork(inpt_dim=inpt_dim, outp_dim=self.model_dim, ctxt_dim=self.ctxt_out, **node_embd_config)
        self.outp_embd = DenseNetwork(inpt_dim=self.model_dim, outp_dim=outp_dim, ctxt_dim=self.ctxt_out, **outp_embd_config)
        self._finish(inpt_dim, outp_dim, ctxt_dim, None)


class FullCrossAttentionEncoder(_DroidNet):
    """droid_transformer.py:622-711."""
    kind = 1

    def __init__(self, inpt_dim: int, outp_dim: int, ctxt_dim: int = 0, cae_config: Mapping = None, node_embd_config: Mapping = None,
                 outp_embd_config: Mapping = None, ctxt_embd_config: Mapping = None):
        super().__init__()
        if not ctxt_dim:
            raise NotImplementedError("the CUDA path expects the time code in the context (ctxt_dim = cond + 2*frequencies)")
        cae_config, node_embd_config = deepcopy(dict(cae_config or {})), deepcopy(dict(node_embd_config or {}))
        outp_embd_config, ctxt_embd_config = deepcopy(dict(outp_embd_config or {})), deepcopy(dict(ctxt_embd_config or {}))
        cae_config["dense_config"] = dict(cae_config.get("dense_config") or {})
        if "model_dim" in cae_config:
            md = cae_config["model_dim"]
            for cfg in (node_embd_config, ctxt_embd_config, outp_embd_config, cae_config["dense_config"]):
                cfg.setdefault("hddn_dim", 2 * md)
        self.ctxt_emdb = DenseNetwork(inpt_dim=ctxt_dim, **ctxt_embd_config)
        self.ctxt_out = self.ctxt_emdb.outp_dim
        self.cae = CrossAttentionEncoder(**cae_config, ctxt_dim=self.ctxt_out)
        self.model_dim = self.cae.model_dim
        self.node_embd = DenseNetwork(inpt_dim=inpt_dim, outp_dim=self.model_dim, ctxt_dim=self.ctxt_out, **node_embd_config)
        self.outp_embd = DenseNetwork(inpt_dim=self.model_dim, outp_dim=outp_dim, ctxt_dim=self.ctxt_out, **outp_embd_config)
        self._finish(inpt_dim, outp_dim, ctxt_dim, None)
