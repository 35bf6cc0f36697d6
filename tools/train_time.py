"""Time the training step (B jets of JetNet-150 per step) with the tensor-core and the CUDA-core kernels (GPU box)."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench

dev = torch.device("cuda:0")
torch.manual_seed(12345)
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
model = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["auto", "cuda_cores", "auto"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
for mode in modes:
    r = bench.train_bench(model, dev, 1, 0, steps=steps, warmup=3 if steps > 2 else 2, B=B, mode=mode)
    print(mode, json.dumps({k: r[k] for k in ("value", "ms_per_step", "final_loss")}))
eng = model.flows[0].net.engine()
print("launches of the last step's library calls:", eng.last_launches())
