"""Host-side handle around the C ABI: owns one ``pfm_epic`` per (module, device) and feeds it torch
CUDA tensors as raw device pointers.  PyTorch is plumbing here (device memory + streams)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib

Tensor = torch.Tensor


@dataclass(frozen=True)
class EpicDims:
    """Resolved EPiC_encoder constructor arguments (epic.py:226-243)."""
    feats: int
    input_dim: int
    hid: int
    latent: int
    layers: int
    t_dim: int
    t_local_cat: bool
    t_global_cat: bool
    global_cond_dim: int = 0
    local_cond_dim: int = 0
    sum_scale: float = 1e-2
    neg_slope: float = 0.01

    @property
    def cond_dim(self) -> int:
        return max(self.global_cond_dim, self.local_cond_dim)

    @property
    def takes_time(self) -> bool:
        return self.t_dim > 0 and (self.t_local_cat or self.t_global_cat)


# Parameter updates that bypass torch's version counters (raw-pointer kernels such as the fused optimizer step) announce
# themselves here; the modules' "did the weights change" key includes this counter.
_GENERATION = [0]


def bump_weights_generation():
    _GENERATION[0] += 1


def weights_generation() -> int:
    return _GENERATION[0]


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: Tensor, device) -> Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class EpicEngine:
    """One packed copy of the network on one GPU."""

    def __init__(self, dims: EpicDims, device: torch.device, precision: str = "fp32"):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.PfmError(f"particle_fm_b200 runs on CUDA devices only (got {device}); there is no CPU fallback")
        self.lib = _lib.load()
        self.dims = dims
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        cfg = _lib.EpicCfgC(dims.feats, dims.input_dim, dims.hid, dims.latent, dims.layers, dims.t_dim,
                            int(dims.t_local_cat), int(dims.t_global_cat), dims.global_cond_dim, dims.local_cond_dim,
                            dims.sum_scale, dims.neg_slope)
        h = C.c_void_p()
        _lib.check(self.lib.pfm_epic_create(C.byref(cfg), self.index, C.byref(h)), "pfm_epic_create")
        self._h = h
        self.n_lin = self.lib.pfm_epic_num_linears(self._h)
        self.precision = "fp32"
        self.set_precision(precision)
        self.weights_key = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self.lib.pfm_epic_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- configuration -------------------------------------------------------------------------
    def set_precision(self, precision: str):
        code = {"fp32": _lib.PFM_PREC_FP32, "bf16": _lib.PFM_PREC_BF16}.get(precision)
        if code is None:
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
        _lib.check(self.lib.pfm_epic_set_precision(self._h, code), "pfm_epic_set_precision")
        self.precision = precision

    def set_train_mode(self, mode: str):
        """'auto' (tensor-core training kernels when hid == 128) or 'cuda_cores' (fused fp32 CUDA-core kernels)."""
        code = {"auto": 0, "cuda_cores": 1}[mode]
        _lib.check(self.lib.pfm_epic_set_train_mode(self._h, code), "pfm_epic_set_train_mode")

    def debug_array(self, which: str, n: int):
        """Host copy (numpy float32) of the first n floats of an internal training array: 'act', 'dact', 'dbeff', 'jact',
        'dpre3' or 'yact' (test hook, pfm_epic_debug_copy)."""
        import numpy as np
        code = {"act": 0, "dact": 1, "dbeff": 2, "jact": 3, "dpre3": 4, "yact": 5}[which]
        buf = np.zeros(n, dtype=np.float32)
        _lib.check(self.lib.pfm_epic_debug_copy(self._h, code, buf.ctypes.data_as(C.c_void_p), n), "pfm_epic_debug_copy")
        return buf

    def linear_shapes(self):
        out = []
        o, i = C.c_int32(), C.c_int32()
        for k in range(self.n_lin):
            _lib.check(self.lib.pfm_epic_linear_shape(self._h, k, C.byref(o), C.byref(i)), "pfm_epic_linear_shape")
            out.append((o.value, i.value))
        return out

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_weights(self, weights: Sequence[Tensor], biases: Sequence[Tensor], key=None):
        """weights[i]: folded [out,in] fp32; biases[i]: [out]; order = state_dict order of the linears."""
        if len(weights) != self.n_lin or len(biases) != self.n_lin:
            raise ValueError(f"expected {self.n_lin} linears, got {len(weights)}/{len(biases)}")
        ws = [_f32c(w, self.device) for w in weights]
        bs = [_f32c(b, self.device) for b in biases]
        for k, ((o, i), w, b) in enumerate(zip(self.linear_shapes(), ws, bs)):
            if tuple(w.shape) != (o, i) or tuple(b.shape) != (o,):
                raise ValueError(f"linear {k}: expected weight {(o, i)} / bias {(o,)}, got {tuple(w.shape)} / {tuple(b.shape)}")
        wp = (C.c_void_p * self.n_lin)(*[w.data_ptr() for w in ws])
        bp = (C.c_void_p * self.n_lin)(*[b.data_ptr() for b in bs])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_set_weights(self._h, wp, bp, self.n_lin, self._stream()), "pfm_epic_set_weights")
        self._keepalive = (ws, bs)      # until the stream has consumed them
        self.weights_key = key

    def _raw_ptrs(self, linears):
        """(key, arrays of device pointers (weight_v | weight, weight_g | NULL, bias), keep-alive tensors) of the linears'
        raw parameters.  The ctypes arrays are cached: optimizers update parameters in place, so between two training steps
        the addresses do not change and the 3 x n_lin tensor conversions are skipped (host time is what bounds a step)."""
        key = tuple((lin.weight_v if lin.weight_norm else lin.weight).data_ptr() for lin in linears) + \
            tuple(lin.bias.data_ptr() for lin in linears) + tuple(lin.weight_g.data_ptr() if lin.weight_norm else 0 for lin in linears)
        c = getattr(self, "_ptr_cache", None)
        if c is not None and c[0] == key:
            return c
        vs, gs, bs = [], [], []
        for k, ((o, i), lin) in enumerate(zip(self.linear_shapes_cached(), linears)):
            v = _f32c(lin.weight_v if lin.weight_norm else lin.weight, self.device)
            g = _f32c(lin.weight_g, self.device) if lin.weight_norm else None
            b = _f32c(lin.bias, self.device)
            if tuple(v.shape) != (o, i) or tuple(b.shape) != (o,) or (g is not None and g.numel() != o):
                raise ValueError(f"linear {k}: expected weight {(o, i)} / bias {(o,)}, got {tuple(v.shape)} / {tuple(b.shape)}")
            vs.append(v); gs.append(g); bs.append(b)
        n = self.n_lin
        arr = lambda ts: (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in ts])
        # cacheable only if the library reads the parameters' own storage (fp32, contiguous, on this device): no copies were made
        own = all(t is None or t.data_ptr() == src.data_ptr() for ts, srcs in
                  ((vs, [lin.weight_v if lin.weight_norm else lin.weight for lin in linears]),
                   (gs, [lin.weight_g if lin.weight_norm else None for lin in linears]),
                   (bs, [lin.bias for lin in linears])) for t, src in zip(ts, srcs) if t is not None)
        c = (key, arr(vs), arr(gs), arr(bs), (vs, gs, bs))
        self._ptr_cache = c if own else None
        return c

    def set_params(self, linears, key=None):
        """Raw parameters of the linears (weight_v / weight_g / bias, or weight / bias for a plain linear): the library folds
        the weight norm and repacks in one launch (pfm_epic_set_params)."""
        if len(linears) != self.n_lin:
            raise ValueError(f"expected {self.n_lin} linears, got {len(linears)}")
        _, av, ag, ab, keep = self._raw_ptrs(linears)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_set_params(self._h, av, ag, ab, self.n_lin, self._stream()), "pfm_epic_set_params")
        self._keepalive = keep
        self.weights_key = key

    def linear_shapes_cached(self):
        if getattr(self, "_shapes", None) is None:
            self._shapes = self.linear_shapes()
        return self._shapes

    def param_grads(self, flat: Tensor, scale: Optional[Tensor], linears):
        """Flat folded-weight gradient -> gradients of the raw parameters, one launch (pfm_epic_param_grads).
        Returns [(d weight_v | d weight, d weight_g | None, d bias)] in the order of ``linears``: views of ONE buffer laid out
        [dv_0 | dg_0 | db_0 | dv_1 | ...] (``self.last_param_grad_buffer``), which a flat optimizer can consume directly."""
        n = self.n_lin
        _, av, ag, _, keep = self._raw_ptrs(linears)
        vs, gs, _ = keep
        lay = getattr(self, "_pg_layout", None)
        if lay is None or lay[0] != n:
            total, offs = 0, []
            for v, g in zip(vs, gs):
                o_v = total; total += v.numel()
                o_g = None
                if g is not None:
                    o_g = total; total += g.numel()
                o_b = total; total += v.shape[0]
                offs.append((o_v, o_g, o_b, tuple(v.shape), None if g is None else tuple(g.shape)))
            lay = (n, total, offs)
            self._pg_layout = lay
        _, total, offs = lay
        buf = torch.empty(total, device=self.device, dtype=torch.float32)     # one allocation, per-parameter views
        sizes = getattr(self, "_pg_sizes", None)
        if sizes is None:
            sizes = []
            for (o_v, o_g, o_b, shv, shg) in offs:
                sizes.append(shv[0] * shv[1])
                if o_g is not None:
                    sizes.append(shg[0] * (shg[1] if len(shg) > 1 else 1))
                sizes.append(shv[0])
            self._pg_sizes = sizes
        parts = buf.split_with_sizes(sizes)           # one call for all views (host time bounds a training step)
        out, k = [], 0
        for (o_v, o_g, o_b, shv, shg) in offs:
            dv = parts[k].view(shv); k += 1
            dg = None
            if o_g is not None:
                dg = parts[k].view(shg); k += 1
            db = parts[k]; k += 1
            out.append((dv, dg, db))
        base = buf.data_ptr()
        adv = (C.c_void_p * n)(*[base + 4 * o[0] for o in offs])
        adg = (C.c_void_p * n)(*[None if o[1] is None else base + 4 * o[1] for o in offs])
        adb = (C.c_void_p * n)(*[base + 4 * o[2] for o in offs])
        scale = None if scale is None else _f32c(scale, self.device).reshape(1)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_param_grads(self._h, _ptr(flat), _ptr(scale), av, ag, adv, adg, adb, n, self._stream()),
                       "pfm_epic_param_grads")
        self.last_param_grad_buffer = buf
        return out

    # -- hot path ------------------------------------------------------------------------------
    def forward(self, t_code: Optional[Tensor], x: Tensor, mask: Optional[Tensor], cond: Optional[Tensor]) -> Tensor:
        """t_code [1|B, t_dim], x [B,N,input_dim], mask [B,N] (or [B,N,1]), cond [B,C] -> [B,N,feats]."""
        d = self.dims
        B, N = int(x.shape[0]), int(x.shape[1])
        x = _f32c(x, self.device)
        if x.shape[2] != d.input_dim:
            raise ValueError(f"x has {x.shape[2]} columns, the net expects input_dim={d.input_dim}")
        mask = None if mask is None else _f32c(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        t_rows = 1
        if d.takes_time:
            if t_code is None:
                raise ValueError("t_local_cat/t_global_cat is set but no time code was given (epic.py:317-321)")
            t_code = _f32c(t_code, self.device).reshape(-1, d.t_dim)
            t_rows = int(t_code.shape[0])
            if t_rows not in (1, B):
                raise ValueError(f"time code must have 1 or B={B} rows, got {t_rows}")
        else:
            t_code = None
        out = torch.empty(B, N, d.feats, device=self.device, dtype=torch.float32)
        self._ticket = getattr(self, "_ticket", 0) + 1      # the call rewrites the handle's plan: a pending backward() is stale
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_forward(self._h, _ptr(t_code), t_rows, _ptr(x), _ptr(mask), _ptr(cond),
                                                 _ptr(out), B, N, self._stream()), "pfm_epic_forward")
        return out

    def sample(self, z: Tensor, mask: Optional[Tensor], cond: Optional[Tensor], t_codes: Optional[Tensor],
               t_codes_in: Optional[Tensor], dt: Tensor, solver: str) -> Tensor:
        """Integrate in place on a copy of z [B,N,feats] (already masked).  t_codes [n_evals,t_dim]."""
        d = self.dims
        B, N = int(z.shape[0]), int(z.shape[1])
        x = z.detach().to(device=self.device, dtype=torch.float32).contiguous().clone()
        mask = None if mask is None else _f32c(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        code = {"euler": _lib.PFM_SOLVER_EULER, "midpoint": _lib.PFM_SOLVER_MIDPOINT}[solver]
        dt = _f32c(dt, self.device).reshape(-1)
        n_steps = int(dt.numel())
        n_evals = n_steps * (2 if solver == "midpoint" else 1)
        if t_codes is not None:
            t_codes = _f32c(t_codes, self.device).reshape(n_evals, -1)
        if t_codes_in is not None:
            t_codes_in = _f32c(t_codes_in, self.device).reshape(n_evals, -1)
        self._ticket = getattr(self, "_ticket", 0) + 1      # the call rewrites the handle's plan: a pending backward() is stale
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_sample(self._h, _ptr(x), _ptr(mask), _ptr(cond), _ptr(t_codes),
                                                _ptr(t_codes_in), _ptr(dt), code, n_steps, B, N, self._stream()),
                       "pfm_epic_sample")
        return x

    # -- training --------------------------------------------------------------------------------
    def grad_size(self) -> int:
        return int(self.lib.pfm_epic_grad_size(self._h))

    def grad_views(self, flat: Tensor):
        """Per-linear (dW [out,in], db [out]) views of the flat gradient buffer (layout of pfm_b200.h)."""
        views, off = [], 0
        for o, i in self.linear_shapes():
            views.append((flat[off:off + o * i].view(o, i), flat[off + o * i:off + o * i + o]))
            off += o * i + o
        return views

    def grad_chunks(self):
        """[(offset, count)] slices of the flat gradient in the order the last backward completes them."""
        n = int(self.lib.pfm_epic_grad_chunks(self._h))
        off, cnt = C.c_longlong(), C.c_longlong()
        out = []
        for i in range(n):
            _lib.check(self.lib.pfm_epic_grad_chunk_range(self._h, i, C.byref(off), C.byref(cnt)), "pfm_epic_grad_chunk_range")
            out.append((off.value, cnt.value))
        return out

    def stream_wait_grad_chunk(self, i: int, stream: "torch.cuda.Stream"):
        _lib.check(self.lib.pfm_epic_stream_wait_grad_chunk(self._h, i, C.c_void_p(stream.cuda_stream)),
                   "pfm_epic_stream_wait_grad_chunk")

    def loss_fwd_bwd(self, kind: str, x: Tensor, mask: Optional[Tensor], cond: Optional[Tensor], t: Tensor,
                     t_code: Optional[Tensor], t_code_in: Optional[Tensor], n0: Tensor, n1: Optional[Tensor],
                     sigma: float, want_grad: bool = True):
        """Fused flow-matching loss (+ backward).  Returns (loss [1] on the device, flat gradient or None)."""
        d = self.dims
        B, N = int(x.shape[0]), int(x.shape[1])
        code = {"FM-OT": _lib.PFM_LOSS_FM_OT, "CFM": _lib.PFM_LOSS_CFM, "droid": _lib.PFM_LOSS_DROID}[kind]
        x = _f32c(x, self.device)
        if x.shape[2] != d.feats:
            raise ValueError(f"x has {x.shape[2]} features, the net expects {d.feats}")
        mask = None if mask is None else _f32c(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        t = _f32c(t, self.device).reshape(B)
        n0 = _f32c(n0, self.device)
        n1 = None if n1 is None else _f32c(n1, self.device)
        t_code = _f32c(t_code, self.device).reshape(B, d.t_dim) if (d.takes_time and t_code is not None) else None
        if d.takes_time and t_code is None:
            raise ValueError("the net takes a time code but none was given")
        t_in = d.input_dim - d.feats
        t_code_in = _f32c(t_code_in, self.device).reshape(B, t_in) if t_in > 0 else None
        loss = torch.empty(1, device=self.device, dtype=torch.float32)
        flat = torch.empty(self.grad_size(), device=self.device, dtype=torch.float32) if want_grad else None
        self._ticket = getattr(self, "_ticket", 0) + 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_loss_fwd_bwd(self._h, _ptr(x), _ptr(t), _ptr(t_code), _ptr(t_code_in), _ptr(n0),
                                                      _ptr(n1), _ptr(mask), _ptr(cond), code, float(sigma), _ptr(loss),
                                                      _ptr(flat), B, N, self._stream()), "pfm_epic_loss_fwd_bwd")
        return loss, flat

    def forward_train(self, t_code: Optional[Tensor], x: Tensor, mask: Optional[Tensor], cond: Optional[Tensor]):
        """forward() that keeps the activations for ONE later backward(); returns (out, ticket)."""
        d = self.dims
        B, N = int(x.shape[0]), int(x.shape[1])
        x = _f32c(x, self.device)
        if x.shape[2] != d.input_dim:
            raise ValueError(f"x has {x.shape[2]} columns, the net expects input_dim={d.input_dim}")
        mask = None if mask is None else _f32c(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        t_rows = 1
        if d.takes_time:
            if t_code is None:
                raise ValueError("t_local_cat/t_global_cat is set but no time code was given (epic.py:317-321)")
            t_code = _f32c(t_code, self.device).reshape(-1, d.t_dim)
            t_rows = int(t_code.shape[0])
            if t_rows not in (1, B):
                raise ValueError(f"time code must have 1 or B={B} rows, got {t_rows}")
        else:
            t_code = None
        out = torch.empty(B, N, d.feats, device=self.device, dtype=torch.float32)
        self._ticket = getattr(self, "_ticket", 0) + 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_forward_train(self._h, _ptr(t_code), t_rows, _ptr(x), _ptr(mask), _ptr(cond),
                                                       _ptr(out), B, N, self._stream()), "pfm_epic_forward_train")
        return out, self._ticket, (t_code, t_rows, cond, B, N)

    def backward(self, ticket: int, saved, grad_out: Tensor, want_gx: bool, want_gw: bool):
        if ticket != getattr(self, "_ticket", 0):
            raise RuntimeError("the saved activations of this forward were overwritten by a later training forward on "
                               "the same network: libpfm_b200 keeps ONE in-flight forward per handle")
        t_code, t_rows, cond, B, N = saved
        d = self.dims
        grad_out = _f32c(grad_out, self.device)
        gx = torch.empty(B, N, d.input_dim, device=self.device, dtype=torch.float32) if want_gx else None
        flat = torch.empty(self.grad_size(), device=self.device, dtype=torch.float32) if want_gw else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_backward(self._h, _ptr(t_code), t_rows, _ptr(cond), _ptr(grad_out), _ptr(gx),
                                                  _ptr(flat), B, N, self._stream()), "pfm_epic_backward")
        return gx, flat

    def _cond(self, cond: Optional[Tensor], B: int) -> Optional[Tensor]:
        d = self.dims
        if d.cond_dim == 0:
            return None        # unconditional nets ignore whatever cond the datamodule hands over (epic.py:347-350)
        if cond is None:
            raise ValueError(f"global_cond_dim={d.global_cond_dim}, local_cond_dim={d.local_cond_dim} but cond is None "
                             "(epic.py:309-313)")
        cond = _f32c(cond, self.device).reshape(B, -1)
        if cond.shape[1] != d.cond_dim:
            raise ValueError(f"cond has {cond.shape[1]} columns, expected {d.cond_dim}")
        return cond

    def set_timing(self, enable: bool):
        _lib.check(self.lib.pfm_epic_set_timing(self._h, int(enable)), "pfm_epic_set_timing")

    def last_kernel_ms(self) -> float:
        """Device time of the fused network/integrator kernel of the last call (needs set_timing(True))."""
        return float(self.lib.pfm_epic_last_kernel_ms(self._h))

    def last_launches(self) -> int:
        return int(self.lib.pfm_epic_last_launches(self._h))

    def last_groups(self) -> int:
        return int(self.lib.pfm_epic_last_groups(self._h))

    # -- diffusion samplers ---------------------------------------------------------
    def sample_diffusion(self, z: Tensor, mask: Optional[Tensor], cond: Optional[Tensor], t_codes: Optional[Tensor],
                         t_codes_in: Optional[Tensor], coef: Tensor, step_kind: str, solver: str = "euler",
                         dt: Optional[Tensor] = None, noise: Optional[Tensor] = None) -> Tensor:
        """Single-launch DDIM / Euler-Maruyama / probability-flow-ODE sampling (pfm_epic_sample_diffusion).
        coef [n_evals, 4] schedule values per evaluation, noise [n_steps, B, N, feats] for 'em'."""
        B, N = int(z.shape[0]), int(z.shape[1])
        x = z.detach().to(device=self.device, dtype=torch.float32).contiguous().clone()
        mask = None if mask is None else _f32c(mask.reshape(B, N), self.device)
        cond = self._cond(cond, B)
        kind = {"pf_ode": _lib.PFM_STEP_PF_ODE, "ddim": _lib.PFM_STEP_DDIM, "em": _lib.PFM_STEP_EM}[step_kind]
        code = {"euler": _lib.PFM_SOLVER_EULER, "midpoint": _lib.PFM_SOLVER_MIDPOINT}[solver]
        coef = _f32c(coef, self.device).reshape(-1, 4)
        n_evals = int(coef.shape[0])
        n_steps = n_evals // (2 if (step_kind == "pf_ode" and solver == "midpoint") else 1)
        dt = None if dt is None else _f32c(dt, self.device).reshape(-1)
        noise = None if noise is None else _f32c(noise, self.device)
        if noise is not None and tuple(noise.shape) != (n_steps, B, N, self.dims.feats):
            raise ValueError(f"noise must be [n_steps={n_steps}, B={B}, N={N}, feats={self.dims.feats}], got {tuple(noise.shape)}")
        if t_codes is not None:
            t_codes = _f32c(t_codes, self.device).reshape(n_evals, -1)
        if t_codes_in is not None:
            t_codes_in = _f32c(t_codes_in, self.device).reshape(n_evals, -1)
        self._ticket = getattr(self, "_ticket", 0) + 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_epic_sample_diffusion(self._h, _ptr(x), _ptr(mask), _ptr(cond), _ptr(t_codes),
                                                          _ptr(t_codes_in), _ptr(coef), _ptr(noise), _ptr(dt), kind, code,
                                                          n_steps, B, N, self._stream()), "pfm_epic_sample_diffusion")
        return x


# ----------------------------------------------------------------------------------------------
# generate_data post-processing
# ----------------------------------------------------------------------------------------------
def postprocess_into(x: Tensor, mask: Optional[Tensor], out: Tensor, scale=None, shift=None, log_col: int = -1,
                     first_only_col: int = -1):
    """out[...] = post(x) (* mask): inverse normalisation / log-pt / masking of data_generation.py:105-123 in one kernel
    that writes straight into ``out`` -- a CUDA tensor or a PINNED host tensor (no separate device -> host copy).
    x [B,N,F] on a CUDA device, mask [B,N(,1)] or None, scale / shift sequences of F python floats or None."""
    if x.device.type != "cuda":
        raise _lib.PfmError("postprocess_into needs a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    B, N, F = (int(s) for s in x.shape)
    x = x.detach().to(torch.float32).contiguous()
    if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (B, N, F):
        raise ValueError("out must be a contiguous float32 tensor of x's shape")
    if out.device.type == "cpu" and not out.is_pinned():
        raise ValueError("a host output buffer must be pinned (page-locked): the kernel writes into it directly")
    if mask is not None:
        mask = mask.detach().to(device=x.device, dtype=torch.float32).reshape(B, N).contiguous()
    sc = sh = None
    if scale is not None:
        sc = (C.c_float * F)(*[float(v) for v in scale])
        sh = (C.c_float * F)(*[float(v) for v in shift])
    with torch.cuda.device(x.device):
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.pfm_postprocess(_ptr(x), _ptr(mask), C.c_void_p(out.data_ptr()), B, N, F, sc, sh, int(log_col),
                                       int(first_only_col), st),
                   "pfm_postprocess")
    if out.device.type == "cuda":
        return out
    x.record_stream(torch.cuda.current_stream(x.device))
    return out


# ----------------------------------------------------------------------------------------------
# CFM-OT coupling
# ----------------------------------------------------------------------------------------------
def ot_assign(x0: Tensor, x1: Tensor, want_cost: bool = False):
    """sigma [B,N] int32: exact optimal assignment of noise particle i to data particle sigma[i] per jet
    (squared Euclidean cost; = POT's ot.emd plan for uniform marginals, losses.py:171-180)."""
    if x0.device.type != "cuda":
        raise _lib.PfmError("ot_assign needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    B, N, F = (int(s) for s in x0.shape)
    a, b = _f32c(x0, x0.device), _f32c(x1, x0.device)
    sigma = torch.empty(B, N, device=x0.device, dtype=torch.int32)
    cost = torch.empty(B, device=x0.device, dtype=torch.float64) if want_cost else None
    with torch.cuda.device(x0.device):
        st = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
        _lib.check(lib.pfm_ot_assign(_ptr(a), _ptr(b), B, N, F, _ptr(sigma), _ptr(cost), st), "pfm_ot_assign")
    return (sigma, cost) if want_cost else sigma


def ot_gather(x0: Tensor, x1: Tensor, mask: Optional[Tensor], sigma: Tensor, pick: Tensor):
    """(x0[k, pick], x1[k, sigma[pick]], mask[k, sigma[pick]]) -- the resampled pairs of losses.py:183-189."""
    lib = _lib.load()
    B, N, F = (int(s) for s in x0.shape)
    a, b = _f32c(x0, x0.device), _f32c(x1, x0.device)
    m = None if mask is None else _f32c(mask.reshape(B, N), x0.device)
    pick = pick.to(device=x0.device, dtype=torch.int32).contiguous()
    x0p, x1p = torch.empty_like(a), torch.empty_like(b)
    mo = torch.empty(B, N, device=x0.device, dtype=torch.float32)
    with torch.cuda.device(x0.device):
        st = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
        _lib.check(lib.pfm_ot_gather(_ptr(a), _ptr(b), _ptr(m), _ptr(sigma), _ptr(pick), B, N, F, _ptr(x0p), _ptr(x1p), _ptr(mo), st),
                   "pfm_ot_gather")
    return x0p, x1p, mo


# ----------------------------------------------------------------------------------------------
# jet-feature flow
# ----------------------------------------------------------------------------------------------
class MlpFlowEngine:
    """Packed copy of a conditional MLP vector field (components/mlp.py small_cond_MLP_model) on one GPU."""

    def __init__(self, features: int, t_dim: int, cond_dim: int, out_widths, concat, act_flags, activation: str,
                 device: torch.device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.PfmError(f"particle_fm_b200 runs on CUDA devices only (got {device}); there is no CPU fallback")
        if activation not in _lib.PFM_ACT:
            raise NotImplementedError(f"activation={activation!r}: the CUDA path implements {sorted(_lib.PFM_ACT)}")
        self.lib = _lib.load()
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.features, self.t_dim, self.cond_dim, self.n = features, t_dim, cond_dim, len(out_widths)
        cfg = _lib.MlpCfgC(features, t_dim, cond_dim, self.n, _lib.PFM_ACT[activation])
        arr = lambda v: (C.c_int32 * self.n)(*[int(i) for i in v])
        h = C.c_void_p()
        _lib.check(self.lib.pfm_mlp_create(C.byref(cfg), arr(out_widths), arr(concat), arr(act_flags), self.index, C.byref(h)),
                   "pfm_mlp_create")
        self._h = h
        self.weights_key = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self.lib.pfm_mlp_destroy(h)
            except Exception:
                pass
            self._h = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_weights(self, weights: Sequence[Tensor], biases: Sequence[Tensor], key=None):
        ws = [_f32c(w, self.device) for w in weights]
        bs = [_f32c(b, self.device) for b in biases]
        o, i = C.c_int32(), C.c_int32()
        for k, (w, b) in enumerate(zip(ws, bs)):
            _lib.check(self.lib.pfm_mlp_linear_shape(self._h, k, C.byref(o), C.byref(i)), "pfm_mlp_linear_shape")
            if tuple(w.shape) != (o.value, i.value) or tuple(b.shape) != (o.value,):
                raise ValueError(f"linear {k}: expected weight {(o.value, i.value)}, got {tuple(w.shape)}")
        wp = (C.c_void_p * self.n)(*[w.data_ptr() for w in ws])
        bp = (C.c_void_p * self.n)(*[b.data_ptr() for b in bs])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_mlp_set_weights(self._h, wp, bp, self.n, self._stream()), "pfm_mlp_set_weights")
        self._keepalive = (ws, bs)
        self.weights_key = key

    def forward(self, t_code: Tensor, x: Tensor, cond: Optional[Tensor]) -> Tensor:
        B = int(x.shape[0])
        x = _f32c(x, self.device).reshape(B, self.features)
        t_code = _f32c(t_code, self.device).reshape(-1, self.t_dim)
        cond = None if cond is None else _f32c(cond, self.device).reshape(B, self.cond_dim)
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_mlp_forward(self._h, _ptr(t_code), int(t_code.shape[0]), _ptr(x), _ptr(cond), _ptr(out), B,
                                                self._stream()), "pfm_mlp_forward")
        return out

    def sample(self, z: Tensor, cond: Optional[Tensor], t_codes: Tensor, dt: Tensor, solver: str) -> Tensor:
        B = int(z.shape[0])
        x = z.detach().to(device=self.device, dtype=torch.float32).contiguous().clone()
        cond = None if cond is None else _f32c(cond, self.device).reshape(B, self.cond_dim)
        code = {"euler": _lib.PFM_SOLVER_EULER, "midpoint": _lib.PFM_SOLVER_MIDPOINT}[solver]
        dt = _f32c(dt, self.device).reshape(-1)
        t_codes = _f32c(t_codes, self.device).reshape(-1, self.t_dim)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pfm_mlp_sample(self._h, _ptr(x), _ptr(cond), _ptr(t_codes), _ptr(dt), code, int(dt.numel()), B,
                                               self._stream()), "pfm_mlp_sample")
        return x
