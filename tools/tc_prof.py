"""Debug: run one bf16 sample() with the in-kernel phase timers (PFM_TC_PROF=1) and print them."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PFM_TC_PROF"] = "1"
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
torch.manual_seed(12345)
m = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to("cuda:0")
m.set_precision("bf16")
mask, n_real = bench.synth_masks(B, 9999)
if len(sys.argv) > 3 and sys.argv[3] == "allreal":
    mask = torch.ones_like(mask)
z = torch.randn(B, 150, 3) * mask
for _ in range(2):
    out = m.flows[0].decode(z.cuda(), None, mask.cuda(), "midpoint", steps)
torch.cuda.synchronize()
print("groups", m.flows[0].net.engine().last_groups(), "jets", B)
