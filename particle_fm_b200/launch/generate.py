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


NOISE_BLOCK = 4096          # jets per noise block of noise="blocks" (fixed: the result must not depend on the world size)


def block_noise(lo: int, hi: int, N: int, F: int, base_seed: int, out: Optional[Tensor] = None) -> Tensor:
    """Rows [lo, hi) of the request's initial noise in `noise="blocks"` mode: the request is cut into fixed blocks of
    NOISE_BLOCK jets and block b is `randn` from its own CPU generator seeded `base_seed + b`, so a rank only draws the
    blocks that overlap its slice (the CPU generator cannot skip ahead) and the values do not depend on the world size."""
    z = out if out is not None else torch.empty(hi - lo, N, F)
    g = torch.Generator()
    b = lo // NOISE_BLOCK
    while b * NOISE_BLOCK < hi:
        b_lo, b_hi = b * NOISE_BLOCK, (b + 1) * NOISE_BLOCK
        g.manual_seed(base_seed + b)
        if b_lo >= lo and b_hi <= hi:                          # whole block inside the slice: drawn in place
            torch.randn((NOISE_BLOCK, N, F), generator=g, out=z[b_lo - lo:b_hi - lo])
        else:
            blk = torch.randn((NOISE_BLOCK, N, F), generator=g)
            s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
            z[s_lo - lo:s_hi - lo] = blk[s_lo - b_lo:s_hi - b_lo]
        b += 1
    return z


def integrate_and_gather(model, z_l: Tensor, c_l: Optional[Tensor], m_l: Optional[Tensor], per: int,
                         ode_solver: str = "midpoint", ode_steps: int = 200, group=None, gather: str = "rank0",
                         parts=None):
    """Device half of the launcher: this rank's (already masked, device-resident) slice -> fused reverse pass -> the
    only collective of the path, the final gather of `per` rows per rank.  Returns the list of per-rank device
    tensors on the gathering ranks (None elsewhere); `parts` lets a caller reuse the receive buffers."""
    rank, world = _world(group)
    n_l = int(z_l.shape[0])
    with torch.no_grad():
        out_l = model.forward(z_l, cond=c_l, mask=m_l, reverse=True, ode_solver=ode_solver, ode_steps=ode_steps) \
            if n_l > 0 else z_l
    if world == 1:
        return [out_l]
    if n_l != per:                                             # short / empty tail rank: zero-padded, trimmed by the caller
        buf = torch.zeros(per, *z_l.shape[1:], dtype=torch.float32, device=z_l.device)
        buf[:n_l] = out_l
    else:
        buf = out_l.contiguous()
    if gather == "all":
        parts = parts if parts is not None else [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        return parts
    if rank == 0 and parts is None:
        parts = [torch.empty_like(buf) for _ in range(world)]
    dist.gather(buf, parts if rank == 0 else None, dst=dist.get_global_rank(group, 0) if group is not None else 0,
                group=group)
    return parts if rank == 0 else None


def generate_data_sharded(model, num_jet_samples: int, cond: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                          ode_solver: str = "midpoint", ode_steps: int = 200, num_points: Optional[int] = None,
                          features: Optional[int] = None, group=None, gather: str = "rank0",
                          integrate: Optional[Callable[[Tensor, Optional[Tensor], Optional[Tensor]], Tensor]] = None,
                          noise: str = "stream") -> Optional[Tensor]:
    """Generate `num_jet_samples` jets on all ranks of `group`.

    Every rank calls this with the SAME arguments (full `cond` / `mask` on the host) and the same CPU RNG state.
    Returns the (num_jet_samples, N, F) result on the CPU on rank 0 (`gather="rank0"`, None elsewhere) or on every
    rank (`gather="all"`).  `noise="stream"` draws the whole request from the default CPU generator exactly like one
    `sample()` call of the reference would and slices it; `noise="blocks"` draws ONE seed from that generator and then
    only this rank's fixed-size noise blocks (`block_noise`), which keeps the host cost per rank constant as ranks are
    added.  Either way the result does not depend on the number of GPUs.  `integrate(z, cond, mask)` replaces the
    module's fused reverse pass in tests that exercise the sharding logic without a GPU (CPU oracle)."""
    rank, world = _world(group)
    N = num_points if num_points else model.hparams.num_particles
    F = features if features else model.hparams.features
    lo, hi = shard_bounds(num_jet_samples, world, rank)
    on_gpu = integrate is None and torch.device(model.device).type == "cuda"
    if noise == "stream":
        z_l = torch.randn(num_jet_samples, N, F)[lo:hi]        # whole request, CPU generator (sample(): :659-662)
    elif noise == "blocks":
        base_seed = int(torch.randint(0, 2 ** 62, (1,)))       # one draw from the default generator, identical on every rank
        z_l = block_noise(lo, hi, N, F, base_seed, out=torch.empty(hi - lo, N, F, pin_memory=True) if on_gpu else None)
    else:
        raise ValueError(f"noise must be 'stream' or 'blocks', got {noise!r}")
    m_l = None if mask is None else mask[lo:hi]
    c_l = None if cond is None else cond[lo:hi]
    per = (num_jet_samples + world - 1) // world
    if integrate is not None:                                  # host-side integration (tests)
        if m_l is not None:
            z_l = z_l * m_l                                    # :669-671
        out_l = integrate(z_l, c_l, m_l) if hi > lo else torch.empty(0, N, F)
        if world == 1:
            return out_l.cpu()
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
    dev = model.device
    z_d = z_l.to(dev, non_blocking=True)
    m_d = None if m_l is None else m_l.to(dev, non_blocking=True)
    c_d = None if c_l is None else c_l.to(dev, non_blocking=True)
    if m_d is not None:
        z_d = z_d * m_d                                        # :669-671
    parts = integrate_and_gather(model, z_d, c_d, m_d, per, ode_solver, ode_steps, group, gather)
    if parts is None:
        return None
    if world == 1:
        return parts[0].cpu()
    res = torch.empty(world * per, N, F, pin_memory=True)
    for r, p_ in enumerate(parts):
        res[r * per:(r + 1) * per].copy_(p_, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return res[:num_jet_samples]
