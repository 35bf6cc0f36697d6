"""Kernel timeline of data-parallel training steps (run under torchrun on >= 2 GPUs): shows the NCCL all-reduce of the flat
gradient slices running on its side stream WHILE the later weight-gradient launches (xty_tc_kernel) execute.
torch.profiler (CUPTI) records every kernel with its stream and start / end time; rank 0 prints the last step.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_overlap_trace.py > profiles/...txt
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
from particle_fm_b200.optim import FusedClipAdamW
from particle_fm_b200.launch import attach_flat_grad_allreduce

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
torch.manual_seed(12345)
model = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to(dev)
attach_flat_grad_allreduce(model)
B = 1024
mask_h, _ = bench.synth_masks(B, 777 + rank)
x = (5.0 * torch.randn(B, bench.N_PART, bench.FEATS) * mask_h).to(dev)
mask = mask_h.to(dev)
opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5)

def step():
    opt.zero_grad(set_to_none=True)
    loss = model.loss(x, mask=mask, cond=None)
    loss.backward()
    opt.step()

for _ in range(5):
    step()
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    # the last step: from the last tbias / plan kernel on
    starts = [i for i, e in enumerate(ev) if "tbias_rows" in e.name]
    ev = ev[starts[-1]:]
    t0 = ev[0].time_range.start
    def short(n):
        n = n.split("(")[0].replace("void ", "").replace("pfm::", "")
        return n[:44]
    print("# rank 0 of %d, one eager data-parallel training step at %d jets per GPU; times in us from the step's first kernel" % (world, B))
    print("#   start       end     dur  kernel")
    nccl, xty = [], []
    for e in ev:
        s, t = e.time_range.start - t0, e.time_range.end - t0
        name = short(e.name)
        if "nccl" in name.lower(): nccl.append((s, t))
        if "xty_tc" in name: xty.append((s, t))
        if "nccl" in name.lower() or "xty_tc" in name or "wn_bwd" in name or "clip_adamw" in name or "sumsq" in name or "rowlin2" in name and s > xty[0][0] if xty else False:
            print("%9.1f %9.1f %7.1f  %s" % (s, t, t - s, name))
    ov = 0.0
    for a, b in nccl:
        for c, d in xty:
            ov += max(0.0, min(b, d) - max(a, c))
    tot = sum(b - a for a, b in nccl)
    print("# all-reduce kernels: %d, %.1f us in all, of which %.1f us (%.0f %%) run while an xty_tc_kernel is executing" % (len(nccl), tot, ov, 100 * ov / max(tot, 1e-9)))
dist.destroy_process_group()
