"""Graphed vs eager training step (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
from particle_fm_b200.optim import FusedClipAdamW
from particle_fm_b200.launch import GraphedTrainStep

dev = torch.device("cuda:0")
for B in [int(a) for a in sys.argv[1:]] or [1024, 128]:
    torch.manual_seed(12345)
    model = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to(dev)
    mask_h, n_real = bench.synth_masks(B, 777)
    x = (5.0 * torch.randn(B, bench.N_PART, bench.FEATS) * mask_h).to(dev)
    mask = mask_h.to(dev)
    opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5, device_step_count=True)
    step = GraphedTrainStep(model, opt, x, mask)
    losses = [float(step(x, mask)) for _ in range(5)]
    torch.cuda.synchronize()
    n = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        l = step(x, mask)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: graphed step {e0.elapsed_time(e1)/n:.3f} ms  -> {B*n/(e0.elapsed_time(e1)*1e-3):.0f} jets/s   losses {losses[:3]} ... {float(l):.4f}")
