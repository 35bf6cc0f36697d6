"""cProfile of the host side of one eager training step (GPU box)."""
import os, sys, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
from particle_fm_b200.optim import FusedClipAdamW

dev = torch.device("cuda:0")
torch.manual_seed(12345)
model = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to(dev)
B = 128
mask_h, _ = bench.synth_masks(B, 777)
x = (5.0 * torch.randn(B, bench.N_PART, bench.FEATS) * mask_h).to(dev)
mask = mask_h.to(dev)
opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5)
def step():
    opt.zero_grad(set_to_none=True)
    loss = model.loss(x, mask=mask, cond=None)
    loss.backward()
    opt.step()
for _ in range(10): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50): step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr).sort_stats("cumulative")
st.print_stats(45)
