import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
torch.manual_seed(12345)
m = SetFlowMatchingLitModule(optimizer=None, **bench.YAML_NET).to("cuda:0"); m.set_precision("bf16")
B = 16384
mask, n_real = bench.synth_masks(B, 9999)
mp = mask.pin_memory()
for _ in range(2): m.sample(B, mask=mp, ode_solver="midpoint", ode_steps=200).cpu()
torch.cuda.synchronize()
t0 = time.perf_counter(); z = torch.randn(B, 150, 3); t1 = time.perf_counter()
print("randn whole %.1f ms" % ((t1 - t0) * 1e3), "threads", torch.get_num_threads())
buf = torch.empty(B * 450, pin_memory=True)
t0 = time.perf_counter(); torch.randn((B, 150, 3), out=buf.view(B, 150, 3)); t1 = time.perf_counter()
print("randn into pinned %.1f ms" % ((t1 - t0) * 1e3))
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x = m.sample(B, mask=mp, ode_solver="midpoint", ode_steps=200)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    y = x.cpu(); t3 = time.perf_counter()
    print("sample() host %.1f ms, +device wait %.1f ms, .cpu() %.1f ms, total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
res = torch.empty(B, 150, 3).pin_memory()
for K in (1, 3, 6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(K):
        res.copy_(m.sample(B, mask=mp, ode_solver="midpoint", ode_steps=200), non_blocking=True)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("K=%d pipelined: host %.1f ms, total %.1f ms = %.1f ms/step" % (K, (t1 - t0) * 1e3, (t2 - t0) * 1e3, (t2 - t0) * 1e3 / K))
