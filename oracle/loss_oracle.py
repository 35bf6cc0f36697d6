"""Oracle (test infrastructure): flow-matching losses and sample(), CPU fp32.

Follows particle_fm/models/components/losses.py literally but with the random draws injectable
so that the CUDA path and the oracle can be fed the same (t, noise):
  FlowMatchingLoss.forward            losses.py:38-77     ("FM-OT", YAML default)
  ConditionalFlowMatchingLoss.forward losses.py:101-136   ("CFM")
  DroidLoss.forward                   losses.py:308-342   ("droid")
and SetFlowMatchingLitModule.sample  flow_matching_module.py:637-677.
Pinned against the reference's own losses.py by oracle/make_golden.py (same CPU RNG stream).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ode_oracle

Tensor = torch.Tensor


def draw_loss_randoms(kind: str, x: Tensor):
    """Consume the default generators in the reference's order.
    t: ``torch.rand_like(torch.ones(B))`` is always on the CPU generator (losses.py:46,104,311);
    noise: ``torch.randn_like(x)`` on x's device (:53 / :108,:116 / :318)."""
    t = torch.rand_like(torch.ones(x.shape[0]))
    n0 = torch.randn_like(x)
    n1 = torch.randn_like(x) if kind == "CFM" else None
    return t, n0, n1


def interpolate(kind: str, x: Tensor, mask: Tensor, t: Tensor, n0: Tensor, n1: Optional[Tensor],
                sigma: float):
    """(y, u_t) for the three loss kinds; t is per jet (B,) and broadcast over particles
    (repeat_interleave in the reference, :47,:105,:312)."""
    tt = t.unsqueeze(-1).repeat_interleave(x.shape[1], dim=1).unsqueeze(-1).type_as(x)
    if kind == "FM-OT":
        y = (1 - tt) * x + (sigma + (1 - sigma) * tt) * n0       # :56
        u = ((1 - sigma) * n0 - x) * mask                        # :61-62
    elif kind == "CFM":
        mu = (1 - tt) * x + tt * n0                              # :115
        y = mu + sigma * n1                                      # :116
        u = (n0 - x) * mask                                      # :118-119
    elif kind == "droid":
        y = x + tt * n0                                          # :320
        u = n0 * mask                                            # :326
    else:
        raise NotImplementedError(kind)
    return tt, y, u


def fm_loss(vf: Callable[[Tensor, Tensor], Tensor], kind: str, x: Tensor, mask: Optional[Tensor],
            t: Tensor, n0: Tensor, n1: Optional[Tensor] = None, sigma: float = 1e-4) -> Tensor:
    """sum((v - u)^2) / sum(mask)  (:75-76, :130, :340-341).  ``vf(t_(B,N), y)`` is the CNF forward."""
    if mask is None:
        if kind == "CFM":
            raise TypeError("ConditionalFlowMatchingLoss needs a mask (losses.py:119,130)")
        mask = torch.ones_like(x[..., 0]).unsqueeze(-1)          # :41-42
    tt, y, u = interpolate(kind, x, mask, t, n0, n1, sigma)
    v = vf(tt.squeeze(-1), y)
    return (v - u).square().sum() / mask.sum()


def sample(vf: Callable[[Tensor, Tensor], Tensor], z: Tensor, mask: Optional[Tensor], ode_solver: str,
           ode_steps: int) -> Tensor:
    """flow_matching_module.py:637-677 after the noise draw: z*mask, integrate 1 -> 0, return traj[-1]."""
    if mask is not None:
        z = z * mask
    return ode_oracle.integrate(vf, z, ode_steps, ode_solver)
