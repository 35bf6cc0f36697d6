"""Throughput of BASELINE.json's other EPiC configurations at their full layer sizes (third argument: fp32 or bf16; the tcgen05 kernels are
specialised for H = 128): generation (midpoint, extrapolated from a short run to the configured ode_steps) and training.
python tools/config_bench.py [c1|c3|c5u|c5c] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule

CFG = {   # name: N, F, H, Z, L, cond (global, local), solver, ode_steps of the shipped config, batch
    "c1": dict(N=30, F=3, H=128, Z=10, L=6, cg=0, cl=0, solver="euler", steps=100),
    "c3": dict(N=279, F=3, H=150, Z=256, L=8, cg=4, cl=4, solver="midpoint", steps=50),
    "c5u": dict(N=128, F=8, H=128, Z=10, L=6, cg=0, cl=0, solver="midpoint", steps=200),
    "c5c": dict(N=128, F=13, H=300, Z=16, L=20, cg=12, cl=0, solver="midpoint", steps=200),
}
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
c = CFG[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
torch.manual_seed(12345)
m = SetFlowMatchingLitModule(optimizer=None, features=c["F"], hidden_dim=c["H"], num_particles=c["N"], frequencies=16, layers=c["L"],
                             latent=c["Z"], t_emb="cosine", t_local_cat=True, t_global_cat=True, add_time_to_input=False,
                             global_cond_dim=c["cg"], local_cond_dim=c["cl"]).to("cuda:0")
m.set_precision(prec)
g = torch.Generator().manual_seed(9999)
n = torch.randint(max(1, c["N"] // 10), c["N"] + 1, (B,), generator=g)
mask = (torch.arange(c["N"]).unsqueeze(0) < n.unsqueeze(1)).float().unsqueeze(-1)
z = (torch.randn(B, c["N"], c["F"], generator=g) * mask).cuda()
cond = torch.randn(B, max(c["cg"], c["cl"]), generator=g).cuda() if max(c["cg"], c["cl"]) else None
mk = mask.cuda()
cnf = m.flows[0]
short = 6 if c["solver"] == "midpoint" else 11
nfe = lambda s: (2 if c["solver"] == "midpoint" else 1) * (s - 1)
for _ in range(2):
    cnf.decode(z, cond, mk, c["solver"], short)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); cnf.decode(z, cond, mk, c["solver"], short); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
per_eval = ms / nfe(short)
gen = B / (per_eval * nfe(c["steps"]) * 1e-3)
# training
x = (5.0 * torch.randn(B, c["N"], c["F"], generator=g) * mask).cuda()
from particle_fm_b200.optim import FusedClipAdamW
from particle_fm_b200.launch import GraphedTrainStep
opt = FusedClipAdamW(m.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5, device_step_count=True)
_graphed = None
def step():          # fused clip + AdamW, the whole step replayed from a CUDA graph (as bench.py's training leg)
    global _graphed
    if _graphed is None:
        _graphed = GraphedTrainStep(m, opt, x, mk, cond)
    _graphed(x, mk, cond)
if B <= 8192:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        step()
    e1.record(); torch.cuda.synchronize()
    tms = e0.elapsed_time(e1) / 5
else:
    tms = float("nan")          # the training call takes at most 12000 jets (the configs train with 256-2048)
print(f"{name} [{prec}]: N={c['N']} F={c['F']} H={c['H']} Z={c['Z']} L={c['L']} cond={c['cg']}/{c['cl']} B={B} mean multiplicity {float(n.float().mean()):.1f}: "
      f"{per_eval:.3f} ms/evaluation -> {gen:.0f} generated jets/s at {c['solver']} ode_steps={c['steps']}; training step {tms:.2f} ms = {B / tms * 1e3:.0f} jets/s")
