import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
torch.manual_seed(12345)
m = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to("cuda:0"); m.set_precision("bf16")
for seed in (9999, 10000, 10001, 10007):
    B = 16384
    mask, n_real = bench.synth_masks(B, seed)
    z = (torch.randn(B, 150, 3) * mask).cuda()
    out = m.flows[0].decode(z, None, mask.cuda(), "midpoint", 3)
    torch.cuda.synchronize()
    g = m.flows[0].net.engine().last_groups()
    rows = int(n_real.sum())
    print(seed, "rows", rows, "min groups", rows / 256, "groups", g, "fill %.4f" % (rows / 256 / g), "waves %.3f" % (g / 148))
