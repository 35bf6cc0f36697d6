"""Is the training step host-bound or device-bound?  Host issue time per step vs device time per step (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule

dev = torch.device("cuda:0")
torch.manual_seed(12345)
model = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mask_h, n_real = bench.synth_masks(B, 777)
x = (5.0 * torch.randn(B, bench.N_PART, bench.FEATS) * mask_h).to(dev)
mask = mask_h.to(dev)
fused = len(sys.argv) > 2 and sys.argv[2] == "fused"
if fused:
    from particle_fm_b200.optim import FusedClipAdamW
    opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5)
else:
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=5e-5)

def step(parts=None):
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    loss = model.loss(x, mask=mask, cond=None)
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    if not fused:
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    if parts is not None:
        parts.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for _ in range(5):
    step()
torch.cuda.synchronize()
parts = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
t0 = time.perf_counter(); e0.record()
for _ in range(n):
    step(parts)
e1.record(); t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
import numpy as np
p = np.array(parts).mean(0) * 1e3
print(f"host issue {1e3*t_issue/n:.3f} ms/step, wall {1e3*t_all/n:.3f} ms/step, device span {e0.elapsed_time(e1)/n:.3f} ms/step")
print(f"host parts (ms): loss fwd+bwd kernels {p[0]:.3f}  backward(param grads) {p[1]:.3f}  clip {p[2]:.3f}  adamw {p[3]:.3f}")
