"""Flow-matching losses -- host-side mirror of particle_fm/models/components/losses.py.

``FlowMatchingLoss`` (:16-77, "FM-OT"), ``ConditionalFlowMatchingLoss`` (:80-136, "CFM") and
``DroidLoss`` (:288-342) keep the reference's constructor, its RNG placement (t on the CPU default
generator, noise on x's device, in the reference's draw order) and return the same scalar with an
autograd graph to the parameters.  The interpolation, the network evaluation, the masked squared
error and the backward pass run in libpfm_b200.so (``particle_fm_b200.training``).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

Tensor = torch.Tensor


class _FusedFMLoss(nn.Module):
    kind: str = ""
    needs_mask: bool = False

    def __init__(self, flows: nn.ModuleList, sigma: float = 1e-4, criterion: str = "mse"):
        super().__init__()
        self.flows = flows
        self.sigma = sigma
        if criterion == "mse":
            self.criterion = nn.MSELoss(reduction="sum")
        elif criterion == "huber":
            if self.kind == "CFM":      # the only loss whose value depends on the criterion (losses.py:130)
                raise NotImplementedError("criterion='huber' with the CFM loss is not supported by the CUDA path")
            self.criterion = nn.HuberLoss(reduction="sum")   # FM-OT / droid ignore it (losses.py:74-76, :339-341)
        else:
            raise NotImplementedError(f"criterion {criterion} not supported")

    def draw(self, x: Tensor):
        """Random draws in the reference's order: t ~ U(0,1) per jet on the CPU generator
        (``torch.rand_like(torch.ones(B))``, losses.py:46,104,311), then noise on x's device."""
        t = torch.rand_like(torch.ones(x.shape[0])).type_as(x)      # .type_as moves it to x's device
        n0 = torch.randn_like(x)
        n1 = torch.randn_like(x) if self.kind == "CFM" else None
        return t, n0, n1

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, cond: Optional[Tensor] = None) -> Tensor:
        if len(self.flows) != 1:
            raise NotImplementedError("n_transforms != 1 is not supported by the CUDA path (1 in every config)")
        if x.dim() != 3:
            raise NotImplementedError("the CUDA path handles set data (B, N, F) only")
        if mask is None:
            if self.needs_mask:
                raise TypeError("ConditionalFlowMatchingLoss needs a mask (the reference fails on mask=None too, "
                                "losses.py:119,130)")
            mask = torch.ones_like(x[..., 0]).unsqueeze(-1)
        t, n0, n1 = self.draw(x)
        from .droid_transformer import _DroidNet, droid_loss_autograd
        if isinstance(self.flows[0].net, _DroidNet):
            return droid_loss_autograd(self.flows[0], self.kind, x, mask, cond, t, n0, n1, float(self.sigma))
        from ...training import fm_loss_autograd
        return fm_loss_autograd(self.flows[0], self.kind, x, mask, cond, t, n0, n1, float(self.sigma))


class FlowMatchingLoss(_FusedFMLoss):
    """y = (1-t) x + (sigma + (1-sigma) t) z;  u = ((1-sigma) z - x) mask;  sum((v-u)^2)/sum(mask)."""
    kind = "FM-OT"


class ConditionalFlowMatchingLoss(_FusedFMLoss):
    """y = (1-t) x1 + t x0 + sigma eps;  u = (x0 - x1) mask;  MSE_sum(v,u)/sum(mask)."""
    kind = "CFM"
    needs_mask = True


class DroidLoss(_FusedFMLoss):
    """y = x + t z;  u = z mask;  sum((v-u)^2)/sum(mask)."""
    kind = "droid"


class ConditionalFlowMatchingOTLoss(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("loss_type='CFM-OT' (POT mini-batch coupling, losses.py:140-204) is a 'next' row of "
                                  "the scope table and not built yet")


class DiffusionLoss(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("loss_type='diffusion' is out of scope of the B200 hot path (SURVEY 2, row 7)")
