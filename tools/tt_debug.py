"""Debug: per-parameter gradient error of the fused training step against the oracle autograd (GPU box)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from helpers import Golden, build_module, rel_l2
from oracle import epic_oracle as eo, loss_oracle as lo

name = sys.argv[1] if len(sys.argv) > 1 else "c1_jetnet30"
kind = sys.argv[2] if len(sys.argv) > 2 else "FM-OT"
g = Golden(name)
m = build_module(g.ctor, g.sd, loss_type=kind, device="cuda:0")
gen = torch.Generator().manual_seed(31)
x = g.x * 5.0 * g.mask
B = x.shape[0]
t = torch.rand(B, generator=gen); n0 = torch.randn(x.shape, generator=gen)
n1 = torch.randn(x.shape, generator=gen) if kind == "CFM" else None
sd = {k: v.clone().requires_grad_(True) for k, v in g.sd.items()}
vf = lambda tt, y: eo.cnf_forward(sd, g.cfg, tt, y, g.cond, g.mask, **g.oracle_kwargs())
ref = lo.fm_loss(vf, kind, x, g.mask, t, n0, n1, 1e-4); ref.backward()
from particle_fm_b200.training import fm_loss_autograd
c = None if g.cond is None else g.cond.cuda()
loss = fm_loss_autograd(m.flows[0], kind, x.cuda(), g.mask.cuda(), c, t.cuda(), n0.cuda(), None if n1 is None else n1.cuda(), 1e-4)
loss.backward()
print("loss", float(loss), float(ref))
for k, p in m.named_parameters():
    if not k.startswith("flows.0.net."): continue
    kk = k[len("flows.0.net."):]
    print(f"{kk:40s} {rel_l2(p.grad.cpu(), sd[kk].grad):.3e}  |ref|={float(sd[kk].grad.norm()):.3e}")
