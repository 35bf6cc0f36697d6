"""Fused clip + AdamW (+ EMA) over flat buffers -- ``pfm_clip_adamw`` behind the torch.optim.Optimizer interface.

The reference steps ``torch.optim.AdamW`` (configs/model/flow_matching.yaml:3-7) after Lightning's
``gradient_clip_val: 0.5`` global-norm clipping and an EMA callback (callbacks/ema.py:73-81): ~20 multi-tensor launches and
~1.2 ms of host time per step for the 87 parameter tensors of the default net -- a third of a training step once the network
itself runs on the tensor cores.  Here the parameters are re-pointed into ONE flat buffer laid out like the flat gradient the
library produces (``EpicEngine.param_grads``), so the whole update is two launches and no per-tensor host work.

Hydra: ``optimizer: {_target_: particle_fm_b200.optim.FusedClipAdamW, _partial_: true, lr: 1e-3, weight_decay: 5e-5,
max_grad_norm: 0.5}`` (and ``trainer.gradient_clip_val: null``: the clipping is inside the step).  Same arithmetic as
``clip_grad_norm_`` + ``AdamW`` (tests/test_gpu_train.py::test_fused_clip_adamw_matches_torch)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch

from . import _lib


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None, ema_decay: Optional[float] = None,
                 device_step_count: bool = False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedClipAdamW takes one parameter group")
        self.max_grad_norm = max_grad_norm
        self.ema_decay = ema_decay
        self.device_step_count = device_step_count     # keep the step count on the device (replay of step() from a CUDA graph)
        self._flat = None          # (params, offsets, flat_p, m, v, ema, ws, ptr_key)
        self._step = 0
        self.last_grad_norm = None

    # -- flat layout ----------------------------------------------------------------------------
    def _layout_from_grads(self, ps):
        """Offsets (floats) of every parameter inside the single buffer all gradients are views of, or None."""
        g0 = ps[0].grad
        base_store = g0.untyped_storage().data_ptr()
        offs, end = [], 0
        for p in ps:
            g = p.grad
            if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.untyped_storage().data_ptr() != base_store:
                return None
            o = (g.data_ptr() - base_store) // 4
            offs.append(o)
            end = max(end, o + g.numel())
        total = sum(p.numel() for p in ps)
        if end != total or len(set(offs)) != len(offs):
            return None
        return offs, base_store, total

    def _layout_fast(self, ps):
        """Steady state: the gradients sit at the offsets adopted at the first step (one data_ptr() per parameter
        instead of the full dtype / contiguity / storage scan)."""
        f = self._flat
        if f is None or len(f[0]) != len(ps):
            return None
        offs, total = f[1], f[2].numel()
        g0 = ps[0].grad
        store = g0.untyped_storage()
        base = store.data_ptr()
        if g0.dtype != torch.float32 or store.nbytes() < 4 * total:
            return None
        for p, o in zip(ps, offs):
            g = p.grad
            if g.data_ptr() != base + 4 * o or g.numel() != p.numel() or not g.is_contiguous():
                return None
        return offs, base, total

    def _flatten(self, ps, offs, total):
        dev = ps[0].device
        flat = torch.empty(total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(ps, offs):
                flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = flat[o:o + p.numel()].view(p.shape)
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        ema = flat.clone() if self.ema_decay is not None else None
        ws = torch.zeros(4, device=dev, dtype=torch.float32)    # [sumsq, grad norm, device-side step count (int), pad]
        key = tuple(p.data_ptr() for p in ps)
        self._flat = (ps, offs, flat, m, v, ema, ws, key)

    def ema_parameters(self):
        """Views of the EMA weights in the order of the optimizer's parameters (None without ema_decay / before the first step)."""
        if self._flat is None or self._flat[5] is None:
            return None
        ps, offs, _, _, _, ema, _, _ = self._flat
        return [ema[o:o + p.numel()].view(p.shape) for p, o in zip(ps, offs)]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        ps = [p for p in group["params"] if p.grad is not None]
        if not ps:
            return loss
        if ps[0].device.type != "cuda":
            raise _lib.PfmError("FusedClipAdamW runs on CUDA devices only (no CPU fallback)")
        lay = self._layout_fast(ps) or self._layout_from_grads(ps)
        if lay is None:              # gradients not produced as views of one buffer: gather them (one multi-tensor copy)
            total = sum(p.numel() for p in ps)
            offs, o = [], 0
            for p in ps:
                offs.append(o); o += p.numel()
            gflat = torch.empty(total, device=ps[0].device, dtype=torch.float32)
            torch._foreach_copy_([gflat[o:o + p.numel()].view(p.shape) for p, o in zip(ps, offs)], [p.grad for p in ps])
        else:
            offs, base_store, total = lay
            g0 = ps[0].grad
            gflat = torch.empty(0, device=g0.device, dtype=torch.float32).set_(g0.untyped_storage(), 0, (total,), (1,))
        f = self._flat
        if f is None or f[1] != offs or f[7] != tuple(p.data_ptr() for p in ps) or len(f[0]) != len(ps):
            if f is not None and f[1] != offs:
                raise RuntimeError("FusedClipAdamW: the gradient layout changed between steps")
            if f is None:
                self._flatten(ps, offs, total)
            else:                      # parameters were moved / re-created (.to(), load with assign): re-adopt them, keep the moments
                _, _, flat, m, v, ema, ws, _ = f
                for p, o in zip(ps, offs):
                    flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
                    p.data = flat[o:o + p.numel()].view(p.shape)
                self._flat = (ps, offs, flat, m, v, ema, ws, tuple(p.data_ptr() for p in ps))
            f = self._flat
        _, _, flat, m, v, ema, ws, _ = f
        self._step += 1
        lib = _lib.load()
        b1, b2 = group["betas"]
        with torch.cuda.device(flat.device):
            st = C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)
            _lib.check(lib.pfm_clip_adamw(C.c_void_p(flat.data_ptr()), C.c_void_p(gflat.data_ptr()), C.c_void_p(m.data_ptr()),
                                          C.c_void_p(v.data_ptr()), total, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                          float(group["weight_decay"]), float(self.max_grad_norm or 0.0),
                                          0 if self.device_step_count else self._step,
                                          None if ema is None else C.c_void_p(ema.data_ptr()), float(self.ema_decay or 0.0),
                                          C.c_void_p(ws.data_ptr()), st), "pfm_clip_adamw")
        gflat.record_stream(torch.cuda.current_stream(flat.device))
        from .engine import bump_weights_generation
        bump_weights_generation()          # the kernel wrote the parameters behind torch's version counters: packed copies are stale
        self.last_grad_norm = ws[1]
        return loss
