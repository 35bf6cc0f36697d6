"""Autograd glue of the fused training kernels.

The reference obtains gradients from torch autograd over the eager network; here the
whole loss forward + backward is libpfm_b200 (``pfm_epic_loss_fwd_bwd``), including the weight-norm fold
``W = g*v/||v||`` (``pfm_epic_set_params``) and its chain rule (``pfm_epic_param_grads``): a
``torch.autograd.Function`` hands autograd the gradients of the raw parameters ``weight_g`` / ``weight_v`` /
``bias``, so optimizers, gradient clipping, DDP hooks and the EMA callback keep working.  (The generic
differentiable forward ``_EpicFn`` keeps the fold as a torch op.)
"""
from __future__ import annotations

from typing import List, Optional

import torch

Tensor = torch.Tensor


def _folded_and_engine(net, device):
    """Folded (W, b) of every linear WITH autograd history, and the engine holding the same values.
    The training kernels are fp32 whatever precision the module uses for sampling."""
    eng = net.engine(device, sync_weights=False)
    folded = [lin.folded() for lin in net.linears()]
    key = net._weights_key()
    if eng.weights_key != key:
        eng.set_weights([w.detach() for w, _ in folded], [b.detach() for _, b in folded], key=key)
    flat: List[Tensor] = []
    for w, b in folded:
        flat += [w, b]
    return eng, flat


def _raw_params(net):
    """[weight_v | weight, weight_g (weight-normed only), bias] of every linear, in the library's order."""
    out = []
    for lin in net.linears():
        out += [lin.weight_v, lin.weight_g, lin.bias] if lin.weight_norm else [lin.weight, lin.bias]
    return out


class _FMLossParamFn(torch.autograd.Function):
    """loss = sum((net(t, y) - u)^2) / sum(mask): forward, backward AND the weight-norm chain rule inside libpfm_b200;
    autograd sees the raw parameters (weight_v, weight_g, bias)."""

    @staticmethod
    def forward(ctx, net, eng, kind, sigma, x, mask, cond, t, t_code, t_code_in, n0, n1, *params):
        want = any(ctx.needs_input_grad[12:])
        loss, flat = eng.loss_fwd_bwd(kind, x, mask, cond, t, t_code, t_code_in, n0, n1, sigma, want_grad=want)
        hook = getattr(net, "flat_grad_hook", None)
        if want and hook is not None:
            flat = hook(flat, eng)                # e.g. the data-parallel all-reduce of the flat gradient
        ctx.net, ctx.eng, ctx.flat = net, eng, flat
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        grads = [None] * 12
        lins = ctx.net.linears()
        if ctx.flat is None:
            return tuple(grads) + (None,) * sum(3 if l.weight_norm else 2 for l in lins)
        for lin, (dv, dg, db) in zip(lins, ctx.eng.param_grads(ctx.flat, g, lins)):
            grads += [dv, dg.view(lin.weight_g.shape), db] if lin.weight_norm else [dv, db]
        return tuple(grads)


_DIRECT_GRADS = False


class direct_param_grads:
    """Context: the fused EPiC losses write ``p.grad`` of the raw parameters themselves (loss forward + backward + weight-norm
    chain rule in the library) and return a detached loss -- no autograd graph, no ``backward()``.  Used by
    launch.GraphedTrainStep: ``backward()`` runs every parameter's AccumulateGrad node on the stream that node was created on,
    and a node kept alive by an earlier eager backward on the default stream (a retained ``loss`` is enough) would pull the
    legacy stream into a CUDA-graph capture (cudaErrorStreamCaptureImplicit)."""

    def __enter__(self):
        global _DIRECT_GRADS
        self._old, _DIRECT_GRADS = _DIRECT_GRADS, True
        return self

    def __exit__(self, *exc):
        global _DIRECT_GRADS
        _DIRECT_GRADS = self._old
        return False


def fm_loss_autograd(cnf, kind: str, x: Tensor, mask: Tensor, cond: Optional[Tensor], t: Tensor, n0: Tensor,
                     n1: Optional[Tensor], sigma: float) -> Tensor:
    """Scalar loss with an autograd graph to ``cnf.net``'s parameters (x: (B,N,F), t: (B,) per jet)."""
    net = cnf.net
    if x.device.type != "cuda":
        raise RuntimeError(f"the flow-matching loss got a batch on {x.device}: the B200 path needs a CUDA device "
                           "(no CPU fallback; use oracle/ for CPU reference numbers)")
    eng = net.engine(x.device)                       # folds the weight norm and repacks when a parameter changed
    takes = net.t_local_cat or net.t_global_cat
    with torch.no_grad():
        code = cnf.time_code(t.to(x.device)) if (takes or cnf.add_time_to_input) else None     # [B, 2*frequencies]
    if _DIRECT_GRADS:
        with torch.no_grad():
            loss, flat = eng.loss_fwd_bwd(kind, x, mask, cond, t, code if takes else None, code if cnf.add_time_to_input else None,
                                          n0, n1, sigma, want_grad=True)
            hook = getattr(net, "flat_grad_hook", None)
            if hook is not None:
                flat = hook(flat, eng)
            lins = net.linears()
            for lin, (dv, dg, db) in zip(lins, eng.param_grads(flat, None, lins)):
                pairs = ((lin.weight_v, dv), (lin.weight_g, dg.view(lin.weight_g.shape)), (lin.bias, db)) if lin.weight_norm \
                    else ((lin.weight, dv), (lin.bias, db))
                for prm, grad in pairs:
                    if prm.requires_grad:
                        prm.grad = grad if prm.grad is None else prm.grad + grad
        return loss.reshape(())
    return _FMLossParamFn.apply(net, eng, kind, sigma, x, mask, cond, t, code if takes else None,
                                code if cnf.add_time_to_input else None, n0, n1, *_raw_params(net))


class _EpicFn(torch.autograd.Function):
    """Differentiable EPiC_encoder.forward: out = net(t_code, x); gradients w.r.t. x and the folded weights."""

    @staticmethod
    def forward(ctx, net, eng, t_code, x, mask, cond, *wb):
        out, ticket, saved = eng.forward_train(t_code, x, mask, cond)
        ctx.eng, ctx.ticket, ctx.saved = eng, ticket, saved
        return out

    @staticmethod
    def backward(ctx, gout):
        want_gx = ctx.needs_input_grad[3]
        want_gw = any(ctx.needs_input_grad[6:])
        gx, flat = ctx.eng.backward(ctx.ticket, ctx.saved, gout, want_gx, want_gw)
        grads = [None, None, None, gx, None, None]
        if flat is None:
            return tuple(grads) + (None,) * (2 * ctx.eng.n_lin)
        for (gw, gb), need_w, need_b in zip(ctx.eng.grad_views(flat), ctx.needs_input_grad[6::2],
                                            ctx.needs_input_grad[7::2]):
            grads.append(gw if need_w else None)
            grads.append(gb if need_b else None)
        return tuple(grads)


def epic_forward_autograd(net, t_code: Optional[Tensor], x_local: Tensor, cond: Optional[Tensor],
                          mask: Optional[Tensor]) -> Tensor:
    if x_local.device.type != "cuda":
        raise RuntimeError(f"EPiC_encoder got an input on {x_local.device}: the B200 path needs a CUDA device "
                           "(no CPU fallback)")
    eng, wb = _folded_and_engine(net, x_local.device)
    return _EpicFn.apply(net, eng, t_code, x_local, mask, cond, *wb)
