"""Throughput of the droid set transformers (BASELINE config 4: JetNet-150, masked attention), fp32 CUDA path.
python tools/droid_bench.py [full|cross] [B] [ode_steps]   -> jets/s measured, and extrapolated to midpoint ode_steps=200."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import bench
from test_droid import NET_CONFIG, MODEL
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule

kind = sys.argv[1] if len(sys.argv) > 1 else "full"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 11
prec = sys.argv[4] if len(sys.argv) > 4 else "fp32"
torch.manual_seed(12345)
m = SetFlowMatchingLitModule(optimizer=None, model=MODEL[kind], features=3, num_particles=150, frequencies=16, t_emb="cosine",
                             add_time_to_input=True, net_config=copy.deepcopy(NET_CONFIG[kind]))
for p in m.parameters():          # the YAML zero-initialises the output layers: randomise so that the field is not 0
    if float(p.abs().max()) == 0:
        torch.nn.init.normal_(p, std=0.02)
m = m.to("cuda:0")
m.set_precision(prec)
mask, n_real = bench.synth_masks(B, 9999)
z = (torch.randn(B, 150, 3) * mask).cuda()
mk = mask.cuda()
cnf = m.flows[0]
for _ in range(2):
    out = cnf.decode(z, None, mk, "midpoint", steps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = cnf.decode(z, None, mk, "midpoint", steps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
nfe = 2 * (steps - 1)
flop_tok = {"full": 4_036_608 + 460_800 * float(n_real.float().mean()) / 150, "cross": 2_608_128}[kind]
tflops = flop_tok * float(n_real.sum()) * nfe / (ms * 1e-3) / 1e12
print(f"{kind} [{prec}]: B={B} mean multiplicity {float(n_real.float().mean()):.1f}, midpoint ode_steps={steps} ({nfe} evaluations): {ms:.1f} ms, "
      f"{ms / nfe:.2f} ms/evaluation, {tflops:.1f} TFLOP/s (as-written FLOPs), launches {cnf.net.engine().last_launches()}; "
      f"extrapolated to ode_steps=200: {B / (ms * 1e-3 * 398 / nfe):.1f} jets/s")
