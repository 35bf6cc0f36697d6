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
