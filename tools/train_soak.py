"""Soak test of the tensor-core training path (GPU box): eager steps with a new random batch shape / mask every step, then
graph replays; checks that the loss stays finite and goes down, and that nothing hangs."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
from particle_fm_b200.optim import FusedClipAdamW
from particle_fm_b200.launch import GraphedTrainStep

dev = torch.device("cuda:0")
torch.manual_seed(1)
model = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to(dev)
opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5)
g = torch.Generator().manual_seed(7)
t0 = time.time()
losses = []
n_eager = int(os.environ.get("SOAK_EAGER", "400"))
for it in range(n_eager):
    B = int(torch.randint(1, 700, (1,), generator=g))
    N = int(torch.randint(1, 151, (1,), generator=g))
    n_real = torch.randint(1, N + 1, (B,), generator=g)
    mask = (torch.arange(N)[None, :] < n_real[:, None]).float().unsqueeze(-1)
    x = (torch.randn(B, N, 3, generator=g) * mask).to(dev)
    opt.zero_grad(set_to_none=True)
    loss = model.loss(x, mask=mask.to(dev), cond=None)
    loss.backward()
    opt.step()
    losses.append(float(loss))
    if it % 50 == 49:                                      # sampling calls between training steps share the engine's buffers
        for prec in ("bf16", "fp32"):
            model.set_precision(prec)
            with torch.no_grad():
                ns = torch.randint(1, bench.N_PART + 1, (33,), generator=g)
                smask = (torch.arange(bench.N_PART)[None, :] < ns[:, None]).float().unsqueeze(-1)
                out = model.sample(33, mask=smask.to(dev), ode_solver="midpoint", ode_steps=5)
            assert bool(torch.isfinite(out).all()), (it, prec)
assert all(l == l and l < 1e4 for l in losses), losses[-5:]
if losses:
    print("eager: %d random-shape steps ok, loss %.3f -> %.3f (%.1f s)" % (n_eager, sum(losses[:20]) / 20, sum(losses[-20:]) / 20, time.time() - t0))
mask_h, _ = bench.synth_masks(1024, 777)
x = (torch.randn(1024, bench.N_PART, bench.FEATS) * mask_h).to(dev)
opt2 = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5, device_step_count=True)
step = GraphedTrainStep(model, opt2, x, mask_h.to(dev))
t0 = time.time()
xm = mask_h.to(dev)
ls = [float(step(x, xm)) if i % 100 == 0 else (step(x, xm), 0.0)[1] for i in range(3000)]
torch.cuda.synchronize()
vals = [l for l in ls if l]
assert all(v == v for v in vals)
print("graph: 3000 replays ok, loss %s (%.1f s)" % (" ".join("%.3f" % v for v in vals[::6]), time.time() - t0))
