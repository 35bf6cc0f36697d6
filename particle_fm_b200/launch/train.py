"""Data-parallel training of the fused flow-matching step.

The reference trains with Lightning DDP (`configs/trainer/ddp.yaml:4-9`): DistributedDataParallel all-reduces the
per-parameter gradients (87 tensors for the default net) in buckets while autograd runs.  Here the fused backward
produces ONE flat fp32 buffer with the gradient of the folded weights (2.25 MB for the default net), so data
parallelism is a single all-reduce (mean) of that buffer over NCCL/NVLink before `torch._weight_norm`'s backward
maps it onto weight_g / weight_v.  The map is linear, so averaging folded-weight gradients equals averaging
parameter gradients -- every rank then takes the identical optimizer step, like DDP.  The loss normaliser stays
per rank (sum(mask) of the local batch), which is DDP's semantics over the reference loss.

Lightning's own DDP keeps working too (the parameters receive ordinary autograd gradients); this hook is the
lighter path for the repo's own launcher and bench.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..training import direct_param_grads


def attach_flat_grad_allreduce(model, group=None):
    """Average the flat gradient over `group` inside the fused training step of every flow of `model`.

    The library finishes the weight gradients in a few consecutive slices of the flat buffer and records an event after
    each; on CUDA the all-reduce of slice i is enqueued on a side stream that waits for that event only, so it overlaps
    with the weight-gradient kernels of the later slices (and nothing on the host blocks).  The main stream waits for all
    slices before the gradient is consumed."""
    world = dist.get_world_size(group)
    side = {}

    def hook(flat: torch.Tensor, eng=None) -> torch.Tensor:
        chunks = eng.grad_chunks() if (eng is not None and flat.is_cuda) else []
        if not chunks:                                    # CPU tensors (gloo tests) or a single slice
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            return flat.div_(world)
        dev = flat.device
        if dev not in side:
            side[dev] = torch.cuda.Stream(device=dev)
        s = side[dev]
        for i, (off, cnt) in enumerate(chunks):
            eng.stream_wait_grad_chunk(i, s)               # slice i of the flat gradient is complete
            with torch.cuda.stream(s):                     # NCCL orders itself after / before the side stream
                piece = flat[off:off + cnt]
                dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=group)
                piece.div_(world)
        torch.cuda.current_stream(dev).wait_stream(s)
        flat.record_stream(s)
        return flat

    for f in model.flows:
        f.net.flat_grad_hook = hook
    return model


def detach_flat_grad_allreduce(model):
    for f in model.flows:
        f.net.flat_grad_hook = None
    return model


class GraphedTrainStep:
    """One training step of the fused flow-matching loss replayed from a CUDA graph.

    A step is ~70 kernel launches plus the autograd / optimizer glue; at small per-GPU batches (strong scaling) the host,
    not the GPU, bounds it.  The whole step -- noise draws on the device, loss forward + backward in libpfm_b200,
    weight-norm chain rule, fused clip + AdamW (``particle_fm_b200.optim.FusedClipAdamW(device_step_count=True)``) -- is
    captured once and replayed; per step the host only draws the per-jet times on the CPU generator (the reference's RNG
    placement, losses.py:46) and copies them plus the batch into static buffers.

        step = GraphedTrainStep(model, optimizer, x_example, mask_example)
        loss = step(x, mask)            # same shapes as the examples

    Library calls are capture-safe after the warm-up steps done here (no allocation, pinned staging for the small tables).
    Data-parallel training: the flat-gradient all-reduce hook is captured with the step when attached before construction."""

    def __init__(self, model, optimizer, x: torch.Tensor, mask: torch.Tensor, cond: torch.Tensor = None, warmup: int = 3):
        if not x.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (no CPU fallback)")
        if getattr(optimizer, "device_step_count", True) is False:
            raise ValueError("FusedClipAdamW must be built with device_step_count=True to be replayed from a graph")
        self.model, self.opt = model, optimizer
        self.x, self.mask = x.clone(), mask.clone()
        self.cond = None if cond is None else cond.clone()
        self.t = torch.zeros(x.shape[0], device=x.device, dtype=torch.float32)
        self._t_pin = torch.empty(x.shape[0], pin_memory=True)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):                     # warm-up on a side stream, as torch.cuda.graphs prescribes
            for _ in range(warmup):
                self._draw_t()
                self._body()
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        self.graph = torch.cuda.CUDAGraph()
        self._draw_t()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()

    def _draw_t(self):
        torch.rand(self._t_pin.shape, out=self._t_pin)    # CPU default generator, like torch.rand_like(torch.ones(B))
        self.t.copy_(self._t_pin, non_blocking=True)

    def _body(self):
        from ..models.components.droid_transformer import _DroidNet
        self.opt.zero_grad(set_to_none=True)
        if any(isinstance(f.net, _DroidNet) for f in self.model.flows):
            loss = self.model.loss(self.x, mask=self.mask, cond=self.cond, t=self.t)
            loss.backward()                               # droid nets: through autograd (see training.direct_param_grads for the pitfall)
        else:
            with direct_param_grads():                    # the library writes p.grad itself: nothing of autograd inside the capture
                loss = self.model.loss(self.x, mask=self.mask, cond=self.cond, t=self.t)
        self.opt.step()
        return loss.detach()

    def __call__(self, x: torch.Tensor, mask: torch.Tensor, cond: torch.Tensor = None) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.mask.copy_(mask, non_blocking=True)
        if cond is not None:
            self.cond.copy_(cond, non_blocking=True)
        self._draw_t()
        self.graph.replay()
        return self.loss
