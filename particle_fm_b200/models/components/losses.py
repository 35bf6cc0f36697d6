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

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, cond: Optional[Tensor] = None, t: Optional[Tensor] = None) -> Tensor:
        """``t`` (optional, (B,) on x's device) replaces the per-jet time draw: the CPU-generator draw cannot be replayed from a
        CUDA graph, so a graphed step (launch.train.GraphedTrainStep) draws it outside and hands it in; the noise draws stay here."""
        if len(self.flows) != 1:
            raise NotImplementedError("n_transforms != 1 is not supported by the CUDA path (1 in every config)")
        if x.dim() != 3:
            raise NotImplementedError("the CUDA path handles set data (B, N, F) only")
        if mask is None:
            if self.needs_mask:
                raise TypeError("ConditionalFlowMatchingLoss needs a mask (the reference fails on mask=None too, "
                                "losses.py:119,130)")
            mask = torch.ones_like(x[..., 0]).unsqueeze(-1)
        if t is None:
            t, n0, n1 = self.draw(x)
        else:
            n0 = torch.randn_like(x)
            n1 = torch.randn_like(x) if self.kind == "CFM" else None
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


class ConditionalFlowMatchingOTLoss(_FusedFMLoss):
    """CFM with mini-batch optimal-transport coupling per jet (losses.py:140-204).

    Reference: for every jet, ``ot.emd`` between the N noise particles and the N (padded) data particles under the
    squared Euclidean cost with uniform marginals (an assignment problem: the plan is a permutation / N), then N pairs
    are drawn from the plan with ``np.random.choice`` and both point sets are re-indexed by the drawn pairs.  Here the
    assignment is ``pfm_ot_assign`` (one warp per jet, exact) and the re-indexing ``pfm_ot_gather``; the pair draws keep
    the reference's RNG (numpy's global RandomState: N uniforms per jet, mapped through the cumulative plan exactly as
    ``RandomState.choice`` does), everything else stays on the device -- no per-jet D2H / POT call.

    Mask: the reference (marked work-in-progress, :139) leaves ``mask_ot`` as the LAST jet's re-indexed mask of shape
    (N, 1) (:189) and hands it to the network, where its own EPiC_encoder raises a shape error (epic.py:370), so the
    reference cannot run this loss end to end.  This implementation keeps every jet's own re-indexed mask (B, N, 1) --
    the evident intent -- for the target and the network, and the reference's normaliser ``mask.sum()`` of the ORIGINAL
    mask (:203).  Coupling, re-indexing and interpolation are pinned against the reference's code
    (oracle/make_golden_next.py).  Like the reference, ``x`` is re-indexed IN PLACE (``x1 = x`` aliases the batch, :152,:188)."""
    kind = "CFM"
    needs_mask = True

    @staticmethod
    def picks_from_uniform(u) -> Tensor:
        """Row picks of ``np.random.choice(N*N, p=plan/plan.sum(), size=N)`` (losses.py:184-185) from its uniform draws u
        [B, N]: the flattened plan has one non-zero (1/N) per row, so choice()'s ``cdf.searchsorted(u, side='right')``
        selects the row whose cumulative mass first exceeds u; the column is that row's assignment."""
        import numpy as np
        N = u.shape[1]
        cdf = np.cumsum(np.full(N, 1.0 / N))
        cdf /= cdf[-1]
        return torch.from_numpy(np.minimum(cdf.searchsorted(u, side="right"), N - 1).astype(np.int32))

    def draw(self, x: Tensor):
        """The reference's draws in its order (:151-193): prior sample on x's device, t on the CPU generator, N uniforms
        per jet from numpy's global RandomState (inside np.random.choice), the sigma noise on the device."""
        import numpy as np
        x0 = torch.randn_like(x)
        t = torch.rand_like(torch.ones(x.shape[0])).type_as(x)
        u = np.random.random_sample((x.shape[0], x.shape[1]))
        eps = torch.randn_like(x0)
        return x0, t, u, eps

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, cond: Optional[Tensor] = None, draws=None) -> Tensor:
        if len(self.flows) != 1:
            raise NotImplementedError("n_transforms != 1 is not supported by the CUDA path (1 in every config)")
        if x.dim() != 3:
            raise NotImplementedError("the CUDA path handles set data (B, N, F) only")
        if mask is None:
            raise TypeError("ConditionalFlowMatchingOTLoss needs a mask (the reference fails on mask=None too, losses.py:189)")
        from ...engine import ot_assign, ot_gather
        x0, t, u, eps = draws if draws is not None else self.draw(x)
        sigma_perm = ot_assign(x0, x)                                                   # :163-180
        x0p, x1p, mask_ot = ot_gather(x0, x, mask, sigma_perm, self.picks_from_uniform(u))   # :183-189
        with torch.no_grad():
            x.copy_(x1p.to(x.dtype))                                                    # the reference re-indexes the batch in place
        m_eff = mask_ot.unsqueeze(-1)
        from ...training import fm_loss_autograd
        loss = fm_loss_autograd(self.flows[0], "CFM", x1p, m_eff, cond, t.to(x.device), x0p, eps, float(self.sigma))
        # the fused kernel normalises by sum(m_eff); the reference by the ORIGINAL mask.sum() (:203)
        return loss * (m_eff.sum() / mask.to(m_eff.dtype).sum())


class DiffusionLoss(nn.Module):
    """Noise-prediction diffusion loss (losses.py:207-285) on the fused network kernels.

    The interpolation, the criterion and the MLE weighting are a handful of elementwise device ops around ONE fused
    network evaluation with autograd (``pfm_epic_forward_train`` / ``pfm_epic_backward`` through ``CNF.forward``)."""

    def __init__(self, flows: nn.ModuleList, sigma: float = 1e-4, criterion: str = "huber",
                 diff_config={"max_sr": 1, "min_sr": 1e-8}):
        super().__init__()
        from .diffusion import VPDiffusionSchedule
        self.flows = flows
        self.sigma = sigma
        self.mle_loss_weight = 0.001
        self.diff_sched = VPDiffusionSchedule(**diff_config)
        if criterion == "mse":
            self.criterion = nn.MSELoss(reduction="none")
        elif criterion == "huber":
            self.criterion = nn.HuberLoss(reduction="none")
        else:
            raise NotImplementedError(f"criterion {criterion} not supported")

    def draw(self, x: Tensor):
        t = torch.rand_like(torch.ones(x.shape[0])).type_as(x)          # :234 CPU generator
        z = torch.randn_like(x)                                         # :240
        return t, z

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, cond: Optional[Tensor] = None, draws=None) -> Tensor:
        if len(self.flows) != 1:
            raise NotImplementedError("n_transforms != 1 is not supported by the CUDA path (1 in every config)")
        if mask is None:
            raise TypeError("DiffusionLoss needs a mask (the reference multiplies by it, losses.py:240)")
        t, z = draws if draws is not None else self.draw(x)
        mask = mask.to(x.dtype)
        noises = z * mask
        times = t.view(-1, 1, 1)
        signal_rates, noise_rates = self.diff_sched(times)              # :255
        noisy_nodes = signal_rates * x + noise_rates * noises           # :258
        tt = t.unsqueeze(-1).expand(-1, x.shape[1])                     # (B, N), what t.squeeze(-1) is in the reference
        pred_noises = self.flows[0](tt, noisy_nodes, mask=mask, cond=cond)
        simple_loss = self.criterion(noises, pred_noises) * mask        # :269
        if self.mle_loss_weight:
            betas = self.diff_sched.get_betas(times)
            mle_loss = (betas / noise_rates) * simple_loss
            return simple_loss.sum() / mask.sum() + self.mle_loss_weight * mle_loss.sum() / mask.sum()
        return simple_loss.sum() / mask.sum()
