"""Debug / profiling: a short bf16 sample() of the bench workload (for ncu captures): tc_small.py [jets] [ode_steps] [allreal]."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
torch.manual_seed(12345)
m = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to("cuda:0")
m.set_precision("bf16")
mask, n_real = bench.synth_masks(B, 9999)
if len(sys.argv) > 3 and sys.argv[3] == "allreal":
    mask = torch.ones_like(mask)
z = torch.randn(B, 150, 3) * mask
eng = m.flows[0].net.engine()
eng.set_timing(True)
for _ in range(2):
    out = m.flows[0].decode(z.cuda(), None, mask.cuda(), "midpoint", steps)
torch.cuda.synchronize()
print("groups", eng.last_groups(), "jets", B, "kernel ms", eng.last_kernel_ms(), "finite", bool(torch.isfinite(out).all()))
