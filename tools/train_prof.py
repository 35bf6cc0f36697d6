"""Debug: a few fused training steps of the default JetNet-150 net (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(12345)
m = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to("cuda:0")
mask, n_real = bench.synth_masks(B, 777)
x = (5.0 * torch.randn(B, 150, 3) * mask).cuda(); mk = mask.cuda()
opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=5e-5)
import time
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    loss = m.loss(x, mask=mk, cond=None)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5); opt.step()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if it == 3: print("groups", m.flows[0].net.engine().last_groups())
    print(f"step {it}: fused fwd+bwd {1e3*(t1-t0):.2f} ms, autograd tail {1e3*(t2-t1):.2f} ms, clip+AdamW {1e3*(t3-t2):.2f} ms")
