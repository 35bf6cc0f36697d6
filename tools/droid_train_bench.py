"""Training-step time of the droid set transformers (fused loss forward + backward through the C ABI), fp32.
python tools/droid_train_bench.py [full|cross] [B]  -> ms per fused step and jets/s (JetNet-150 shape, variable multiplicity)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import bench
from test_droid import NET_CONFIG, MODEL
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
from particle_fm_b200.models.components.droid_transformer import droid_loss_autograd

kind = sys.argv[1] if len(sys.argv) > 1 else "full"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(12345)
m = SetFlowMatchingLitModule(optimizer=None, model=MODEL[kind], features=3, num_particles=150, frequencies=16, t_emb="cosine",
                             add_time_to_input=True, loss_type="droid", net_config=copy.deepcopy(NET_CONFIG[kind]))
for p in m.parameters():
    if float(p.abs().max()) == 0:
        torch.nn.init.normal_(p, std=0.02)
m = m.to("cuda:0")
mask, n_real = bench.synth_masks(B, 9999)
x = (torch.randn(B, 150, 3) * mask).cuda()
mk = mask.cuda()
cnf = m.flows[0]
opt = torch.optim.AdamW(cnf.net.parameters(), lr=1e-4)
t = torch.rand(B).cuda()
n0 = torch.randn(B, 150, 3).cuda()


def step():
    opt.zero_grad(set_to_none=True)
    loss = droid_loss_autograd(cnf, "droid", x, mk, None, t, n0, None, 1e-4)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"{kind}: B={B} (dense rows {B * 150}, mean multiplicity {float(n_real.float().mean()):.1f}): {ms:.1f} ms per training step "
      f"(fused loss fwd+bwd + AdamW), {B / ms * 1e3:.0f} jets/s, launches {cnf.net.engine().last_launches()}, loss {float(loss):.4f}")
