"""Jet-sharded generation over the GPUs of one box.

The reference generates on one device (`utils/data_generation.py:77-123`).  Jets are independent, so the request is
cut into contiguous slices, one per rank; every rank integrates its slice with the fused kernel and the only
collective is the final gather of (n_local, N, F) fp32 to rank 0.  The initial noise is drawn ONCE per request from
the CPU default generator -- the stream `SetFlowMatchingLitModule.sample` uses (flow_matching_module.py:659-663) --
on every rank identically (same seed) and sliced, so the result does not depend on the number of GPUs.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank `rank`: ceil(n / world) jets per rank, the last ranks may be short or empty."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def _world(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def generate_data_sharded(model, num_jet_samples: int, cond: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                          ode_solver: str = "midpoint", ode_steps: int = 200, num_points: Optional[int] = None,
                          features: Optional[int] = None, group=None, gather: str = "rank0",
                          integrate: Optional[Callable[[Tensor, Optional[Tensor], Optional[Tensor]], Tensor]] = None
                          ) -> Optional[Tensor]:
    """Generate `num_jet_samples` jets on all ranks of `group`.

    Every rank calls this with the SAME arguments (full `cond` / `mask` on the host) and the same CPU RNG state.
    Returns the (num_jet_samples, N, F) result on the CPU on rank 0 (`gather="rank0"`, None elsewhere) or on every
    rank (`gather="all"`).  `integrate(z, cond, mask)` defaults to the module's fused reverse pass
    (`model.forward(..., reverse=True)`); tests substitute the CPU oracle to exercise the sharding logic without a GPU.
    """
    rank, world = _world(group)
    N = num_points if num_points else model.hparams.num_particles
    F = features if features else model.hparams.features
    z = torch.randn(num_jet_samples, N, F)                    # whole request, CPU generator (sample(): :659-662)
    lo, hi = shard_bounds(num_jet_samples, world, rank)
    if integrate is None:
        dev = model.device

        def integrate(zl, cl, ml):
            with torch.no_grad():
                return model.forward(zl.to(dev), cond=None if cl is None else cl.to(dev),
                                     mask=None if ml is None else ml.to(dev), reverse=True, ode_solver=ode_solver,
                                     ode_steps=ode_steps)
    z_l = z[lo:hi]
    m_l = None if mask is None else mask[lo:hi]
    c_l = None if cond is None else cond[lo:hi]
    if m_l is not None:
        z_l = z_l * m_l                                        # :669-671
    if hi > lo:
        out_l = integrate(z_l, c_l, m_l)
    else:
        out_l = torch.empty(0, N, F)
    if world == 1:
        return out_l.cpu()
    # final gather: every rank contributes ceil(n/world) rows (short ranks are zero-padded, trimmed after)
    per = (num_jet_samples + world - 1) // world
    backend = dist.get_backend(group)
    dev = out_l.device if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(per, N, F, dtype=torch.float32, device=dev)
    buf[:hi - lo] = out_l.to(dev)
    if gather == "all":
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if rank != 0:
            return None
    return torch.cat(parts)[:num_jet_samples].cpu()
