#!/usr/bin/env python
"""Headline benchmark: EPiC-FM JetNet-150 generation, jets/s (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--precision fp32|bf16] [--batch B]
    python bench.py --impl reference ...          # the CPU port of the reference, same metric/config

A "step" is one full ``sample()`` pass of one batch of B jets per GPU: midpoint, ode_steps=200
(199 steps, 398 network evaluations), random-init default JetNet net (configs/model/flow_matching.yaml),
synthetic variable-multiplicity prefix masks (n_real ~ U[15,150]).  Jets are sharded over the ranks
(no data-path collective; a final gather of the results to rank 0 is inside the timed region for N>1).

value     device-timed throughput with the step's inputs already resident in HBM
e2e       the same through the public API (SetFlowMatchingLitModule.sample(...).cpu()): CPU-generator
          noise, pinned-host -> device copies and the device -> host read of the result inside the timing
roofline  algorithmic FLOPs of the fused network/integrator kernel (real particles only, hoisted form,
          SURVEY 8d) / its CUDA-event duration, against MEASURED_PEAKS.json
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PART, FEATS, ODE_STEPS, SOLVER = 150, 3, 200, "midpoint"
NFE = 2 * (ODE_STEPS - 1)
YAML_NET = dict(features=FEATS, hidden_dim=128, num_particles=N_PART, frequencies=16, layers=6, latent=10,
                t_emb="cosine", t_local_cat=True, t_global_cat=True, add_time_to_input=False)
WORKLOAD = ("EPiC-FM JetNet-150 (150x3, variable-multiplicity masks) midpoint ode_steps=200 generation, "
            "random-init default net (H128 Z10 L6 T32), jets sharded over ranks")
WORKLOAD_ALL_REAL = WORKLOAD.replace("variable-multiplicity masks", "all 150 particles real")
# FLOP per real particle per evaluation and per jet per evaluation (hoisted form)
FLOP_PER_PARTICLE, FLOP_PER_JET = 427_520, 684_096


def synth_masks(B, seed):
    g = torch.Generator().manual_seed(seed)
    n_real = torch.randint(max(1, N_PART // 10), N_PART + 1, (B,), generator=g)
    mask = (torch.arange(N_PART)[None, :] < n_real[:, None]).float().unsqueeze(-1)
    return mask, n_real


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6552.0), tf_burst=d.get("bf16_tflops", 1648.4),
                    tf_sustained=d.get("bf16_tflops_sustained", 1383.1), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_sampler():
    """CPU port of the reference path (oracle/): same net, same seeds, fp32, all host threads."""
    from oracle import epic_oracle as eo, loss_oracle as lo
    cfg = eo.EpicCfg(feats=FEATS, input_dim=FEATS, hid=128, latent=10, layers=6, t_dim=32, t_local_cat=True,
                     t_global_cat=True)
    torch.manual_seed(12345)
    from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
    m = SetFlowMatchingLitModule(optimizer=None, **YAML_NET)            # default init, seed 12345 (fm_tops150.yaml:19)
    sd = {k[len("flows.0.net."):]: v.detach() for k, v in m.state_dict().items() if k.startswith("flows.0.net.")}

    def run(z, mask):
        vf = lambda t, y: eo.cnf_forward(sd, cfg, t, y, None, mask, t_emb="cosine", frequencies=16,
                                         add_time_to_input=False)
        with torch.no_grad():
            return lo.sample(vf, z, mask, SOLVER, ODE_STEPS)
    return run


def eager_gpu_leg(dev, n_jets):
    """The oracle port (the reference's module arithmetic, eager PyTorch ops, fp32) run on the GPU: what a user of the
    reference gets on one B200 without this library.  Outside every timed region of the product path."""
    from oracle import epic_oracle as eo, loss_oracle as lo
    cfg = eo.EpicCfg(feats=FEATS, input_dim=FEATS, hid=128, latent=10, layers=6, t_dim=32, t_local_cat=True, t_global_cat=True)
    sd = {k: v.to(dev) for k, v in eo.synth_state_dict(cfg, 12345).items()}
    mask, n_real = synth_masks(n_jets, 9999)
    z = (torch.randn(n_jets, N_PART, FEATS, generator=torch.Generator().manual_seed(1)) * mask).to(dev)
    mask = mask.to(dev)
    vf = lambda t, y: eo.cnf_forward(sd, cfg, t.to(dev), y, None, mask, t_emb="cosine", frequencies=16, add_time_to_input=False)
    with torch.no_grad():
        lo.sample(vf, z, mask, SOLVER, 4)                        # warm-up (3 steps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lo.sample(vf, z, mask, SOLVER, ODE_STEPS)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": n_jets / dt, "unit": "jets/s", "kind": "oracle port, eager PyTorch fp32 on cuda:0 (one launch per op)",
            "sample": f"{n_jets} jets (the reference's eval batch size, jetnet_eval.yaml:16-19), full 398 evaluations, {dt:.2f} s"}


def time_cpu(run, n_jets, seed, reps=1):
    mask, n_real = synth_masks(n_jets, seed)
    g = torch.Generator().manual_seed(seed + 1)
    z = torch.randn(n_jets, N_PART, FEATS, generator=g)
    t0 = time.perf_counter()
    for _ in range(reps):
        run(z, mask)
    dt = (time.perf_counter() - t0) / reps
    return n_jets / dt, dt, float(n_real.float().mean())


def train_bench(model, dev, world, rank, steps=8, warmup=3, B=1024, mode="auto"):
    """Secondary metric (BASELINE.json: "train jets/s"): full training steps of the default JetNet-150 net --
    fused FM-OT loss forward+backward (fp32 kernels), flat-gradient all-reduce over the ranks, global-norm clip 0.5,
    AdamW(1e-3, wd 5e-5) (configs/model/flow_matching.yaml:3-7, experiment gradient_clip_val 0.5).  Weak scaling:
    B jets per GPU per step (jetnet_tops_30_jedi.yaml:3 batch 1024)."""
    import torch.distributed as dist
    from particle_fm_b200.launch import attach_flat_grad_allreduce
    if world > 1:
        attach_flat_grad_allreduce(model)
    mask_h, n_real = synth_masks(B, 777 + rank)
    g = torch.Generator().manual_seed(888 + rank)
    x = (5.0 * torch.randn(B, N_PART, FEATS, generator=g) * mask_h).to(dev)
    mask = mask_h.to(dev)
    fused_opt = mode == "auto"
    graphed = fused_opt and os.environ.get("PFM_BENCH_NO_GRAPH") is None
    model.flows[0].net.engine().set_train_mode(mode)
    if fused_opt:          # clip 0.5 + AdamW in two launches over flat buffers (particle_fm_b200.optim)
        from particle_fm_b200.optim import FusedClipAdamW
        opt = FusedClipAdamW(model.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5, device_step_count=graphed)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=5e-5)
    if graphed:            # the whole step replayed from a CUDA graph (launch.GraphedTrainStep); t is drawn on the CPU per step
        from particle_fm_b200.launch import GraphedTrainStep
        gstep = GraphedTrainStep(model, opt, x, mask)

    def step():
        if graphed:
            return gstep(x, mask)
        opt.zero_grad(set_to_none=True)
        loss = model.loss(x, mask=mask, cond=None)
        loss.backward()
        if not fused_opt:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    return {"value": world * B * steps / (ms / 1e3), "unit": "jets/s", "batch_per_gpu": B, "steps": steps,
            "ms_per_step": ms / steps, "loss": "FM-OT", "final_loss": float(loss.detach()),
            "step": ("loss fwd+bwd with the per-particle GEMMs and their transposes on tcgen05 (3-term bf16 split, fp32-accurate), per-jet "
                     "MLPs on CUDA cores" if mode == "auto" else "fused loss fwd+bwd on fp32 CUDA cores") +
                    "; weight gradients on tcgen05 (3-term bf16 split) + in-library weight-norm fold/chain rule + flat-grad all-reduce + " +
                    ("clip 0.5 + AdamW fused over flat buffers (pfm_clip_adamw, 2 launches)" if fused_opt else "torch clip_grad_norm_ 0.5 + torch.optim.AdamW") +
                    ("; the whole step replayed from a CUDA graph (per-jet times drawn on the CPU generator each step)" if graphed else ""),
            "kernels": mode, "cuda_graph": bool(graphed),
            "mean_real_particles": float(n_real.float().mean())}


def reference_arm(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port; torchdyn/Lightning are
    not installable here), all host threads, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    run = oracle_sampler()
    n_jets = args.ref_jets
    for _ in range(max(1, min(args.warmup, 1))):
        time_cpu(run, n_jets, 9999)
    t0 = time.perf_counter()
    for s in range(args.steps):
        time_cpu(run, n_jets, 9999 + s)
    dt = time.perf_counter() - t0
    value = n_jets * args.steps / dt
    sample = f"{n_jets} jets per step, full midpoint ode_steps=200 (398 evaluations), mean multiplicity ~82"
    line = {"impl": "reference", "metric": "generated_jets_per_s", "value": value, "unit": "jets/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "ode_steps": ODE_STEPS, "solver": SOLVER, "nfe": NFE},
            "cpu_baseline": {"value": value, "unit": "jets/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "jets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PFM_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=0, help="jets per GPU per step (default: 16768 bf16 / 4096 fp32)")
    ap.add_argument("--all-real", action="store_true", help="every particle real (roofline variant)")
    ap.add_argument("--ref-jets", type=int, default=192, help="jets per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-jets", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-throughput measurement")
    ap.add_argument("--no-all-real", action="store_true", help="skip the all-real roofline variant")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank)
    if args.warmup < 3:
        args.warmup = 3                                            # timing rule: >= 3 warm-up steps

    import torch.distributed as dist
    from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # 16768 jets of mean multiplicity ~82.4 bin-pack into ~36.6 x 148 groups of 256 rows: every rank's mask (different seeds,
    # +-0.7 % rows) then needs 37 full-machine waves of the persistent kernel, none tips over into an extra, nearly empty one
    B = args.batch or (16768 if args.precision == "bf16" else 4096)

    torch.manual_seed(12345)                                       # the configs' seed (fm_tops150.yaml:19)
    model = SetFlowMatchingLitModule(optimizer=None, **YAML_NET).to(dev)
    model.set_precision(args.precision)
    cnf = model.flows[0]
    eng = cnf.net.engine()
    eng.set_timing(True)

    # this rank's shard of the request: masks + CPU-generator noise, resident in HBM before the timed region
    mask_h, n_real = synth_masks(B, 9999 + rank)
    if args.all_real:
        mask_h = torch.ones_like(mask_h); n_real = torch.full_like(n_real, N_PART)
    g = torch.Generator().manual_seed(4242 + rank)
    z_h = torch.randn(B, N_PART, FEATS, generator=g) * mask_h
    mask_d, z_d = mask_h.to(dev), z_h.to(dev)
    gather_buf = [torch.empty(B, N_PART, FEATS, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2
    from particle_fm_b200.launch import generate_data_sharded, integrate_and_gather

    def step():
        # the product launcher's device half: fused reverse pass of this rank's slice + the final gather to rank 0
        return integrate_and_gather(model, z_d, None, mask_d, B, SOLVER, ODE_STEPS, parts=gather_buf)

    for _ in range(args.warmup):
        step()
    launches_per_step = eng.last_launches()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    for s in range(args.steps):
        flush.fill_(s & 0xFF)                                      # flush L2 between timed iterations (not timed)
        ev[s][0].record()
        step()
        ev[s][1].record()
        ev[s][1].synchronize()
        kernel_ms.append(eng.last_kernel_ms())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * args.steps / (total_ms / 1e3)

    # ---- multi-GPU correctness, outside the timed region: rank 0 re-integrates every other rank's slice (inputs
    # regenerated from the rank's seeds) on its own GPU and compares with what the gather delivered: bit-equal ----
    multi_check = None
    if world > 1 and rank == 0:
        bad = 0
        for r in range(1, world):
            m_r, _ = synth_masks(B, 9999 + r)
            if args.all_real:
                m_r = torch.ones_like(m_r)
            z_r = torch.randn(B, N_PART, FEATS, generator=torch.Generator().manual_seed(4242 + r)) * m_r
            ref_r = cnf.decode(z_r.to(dev), None, m_r.to(dev), ode_solver=SOLVER, ode_steps=ODE_STEPS)
            bad += int(not torch.equal(ref_r, gather_buf[r]))
        if bad:
            raise SystemExit(f"multi-GPU check failed: {bad} of {world - 1} gathered slices differ from their single-GPU re-integration")
        multi_check = {"slices_checked": world - 1, "bit_equal_to_single_gpu": True,
                       "how": "rank 0 re-integrated every other rank's slice after the timed region and compared with the gathered tensor (torch.equal)"}
    if world > 1:
        dist.barrier()

    # ---- end to end through the public API: CPU noise, H2D of inputs, integration, D2H of the result ----
    e2e_steps = max(1, min(args.steps, 3))
    latency_ms = None
    gen_data = None
    if world == 1:
        mask_pin = mask_h.pin_memory()
        res_pin = torch.empty(B, N_PART, FEATS).pin_memory()                            # the result lands in pinned host memory
        res_pin.copy_(model.sample(B, mask=mask_pin, ode_solver=SOLVER, ode_steps=ODE_STEPS))   # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()                                                        # one isolated call, nothing overlapped
        res_pin.copy_(model.sample(B, mask=mask_pin, ode_solver=SOLVER, ode_steps=ODE_STEPS))
        torch.cuda.synchronize()
        latency_ms = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res_pin.copy_(model.sample(B, mask=mask_pin, ode_solver=SOLVER, ode_steps=ODE_STEPS), non_blocking=True)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_api = ("SetFlowMatchingLitModule.sample(n, mask=pinned) copied to a pinned host buffer (CPU noise draw, H2D, "
                   "integration, D2H inside the timed region); back-to-back calls, the CPU draw of call k+1 overlaps the GPU work of call k")
        # the caller one level up: generate_data (batching + inverse normalisation + masking), post-processing on the device,
        # results written by the kernel into one pinned host buffer
        from particle_fm_b200.utils.data_generation import generate_data
        gd_kw = dict(batch_size=B // 4, device=str(dev), variable_set_sizes=True, mask=mask_pin, normalized_data=True, normalize_sigma=5,
                     means=[0.0, 0.0, 0.02], stds=[0.1, 0.1, 0.03], verbose=False, ode_solver=SOLVER, ode_steps=ODE_STEPS)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gd_out, _ = generate_data(model, B, **gd_kw)
        gd_s = time.perf_counter() - t0
        gen_data = {"value": B / gd_s, "unit": "jets/s", "jets": B, "batch_size": B // 4,
                    "api": "particle_fm_b200.utils.data_generation.generate_data(model, n, batch_size, mask, normalized_data=True, ...): "
                           "4 batches, device post-processing straight into one pinned host buffer, whole call timed on the host clock"}
        assert gd_out.shape == (B, N_PART, FEATS)
    else:
        # N > 1: the product launcher on host inputs -- every rank draws its noise blocks, copies them and its mask slice
        # to its GPU, integrates, the slices are gathered to rank 0 and copied into pinned host memory
        mask_all = torch.cat([synth_masks(B, 9999 + r)[0] for r in range(world)])
        if args.all_real:
            mask_all = torch.ones_like(mask_all)
        mask_all = mask_all.pin_memory()
        torch.manual_seed(777)
        generate_data_sharded(model, world * B, mask=mask_all, ode_solver=SOLVER, ode_steps=ODE_STEPS, noise="blocks")   # warm
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            generate_data_sharded(model, world * B, mask=mask_all, ode_solver=SOLVER, ode_steps=ODE_STEPS, noise="blocks")
        torch.cuda.synchronize()
        dist.barrier()
        e2e_s = time.perf_counter() - t0
        e2e_api = ("particle_fm_b200.launch.generate_data_sharded(model, n, mask=host, noise='blocks'): per-rank CPU noise draw, "
                   "H2D, integration, gather to rank 0, D2H into pinned host memory, all inside the timed region")
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(te.item())
    h2d = B * N_PART * FEATS * 4 + B * N_PART * 4 + NFE * 32 * 4 + (ODE_STEPS - 1) * 4
    d2h = B * N_PART * FEATS * 4

    # ---- all-real variant of the same kernel (the variant the roofline fraction is also quoted on) ----
    all_real = None
    if world == 1 and not args.all_real and args.precision == "bf16" and not args.no_all_real:
        B_ar = 148 * 28
        ones = torch.ones(B_ar, N_PART, 1, device=dev)
        z_ar = torch.randn(B_ar, N_PART, FEATS, generator=torch.Generator().manual_seed(4243)).to(dev)
        ms = []
        for i in range(3):
            flush.fill_(i)
            cnf.decode(z_ar, None, ones, ode_solver=SOLVER, ode_steps=ODE_STEPS)
            torch.cuda.synchronize()
            ms.append(eng.last_kernel_ms())
        k_ar = statistics.mean(ms[1:])
        fl_ar = (B_ar * N_PART * FLOP_PER_PARTICLE + B_ar * FLOP_PER_JET) * NFE
        all_real = {"jets": B_ar, "kernel_ms": k_ar, "jets_per_s": B_ar / (k_ar * 1e-3), "achieved": fl_ar / (k_ar * 1e-3) / 1e12}

    # ---- the reference's own eager PyTorch arithmetic (oracle port) on this GPU: a GPU comparator for the speed-up ----
    gpu_eager = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            gpu_eager = eager_gpu_leg(dev, 1024)
        except Exception as e:                                   # reported, never fatal: it is a comparator only
            gpu_eager = {"error": repr(e)[:200]}

    train = None
    if not args.no_train:
        train = train_bench(model, dev, world, rank)
        if world == 1:         # A/B: the fused fp32 CUDA-core kernels of round 1
            cc = train_bench(model, dev, world, rank, mode="cuda_cores")
            train["cuda_core_kernels"] = {k: cc[k] for k in ("value", "unit", "ms_per_step")}
            model.flows[0].net.engine().set_train_mode("auto")
        if world > 1:          # also the strong-scaling figure, global batch 1024 split over the ranks
            strong = train_bench(model, dev, world, rank, B=max(1, 1024 // world))
            train["strong_scaling"] = {k: strong[k] for k in ("value", "unit", "batch_per_gpu", "ms_per_step")}

    if rank == 0:
        peaks = measured_peaks()
        flops = (float(n_real.sum()) * FLOP_PER_PARTICLE + B * FLOP_PER_JET) * NFE          # per launch, this rank
        k_ms = statistics.mean(kernel_ms)
        achieved = flops / (k_ms * 1e-3) / 1e12
        traffic, traffic_source = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if args.precision == "bf16" and os.path.exists(tpath):          # DRAM bytes of one ncu --set full capture, scaled per jet
            tj = json.load(open(tpath))["epic_tc_kernel"]
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["jets"] * B
            traffic_source = (f"static: one ncu --set full capture ({tj.get('capture', 'profiles/')}, {tj['jets']} jets), "
                              "scaled by jets per launch; not measured in this run")
        peak = peaks["tf_sustained"]
        line = {"metric": "generated_jets_per_s", "value": value, "unit": "jets/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD_ALL_REAL if args.all_real else WORKLOAD, "batch_per_gpu": B, "ode_steps": ODE_STEPS, "solver": SOLVER, "nfe": NFE,
                           "precision": args.precision, "mean_real_particles": float(n_real.float().mean()),
                           "all_real": bool(args.all_real), "l2": "flushed between timed steps (256 MB write)",
                           "parallelism": f"jets sharded x{world}, final gather to rank 0"},
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "jets/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "api": e2e_api, "single_call_latency_ms": latency_ms, "generate_data": gen_data},
                "gpu_launches": launches_per_step * args.steps,
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                             "algorithmic_bytes_per_launch": B * (2 * N_PART * FEATS * 4 + N_PART * 4),
                             "kernel": "epic_tc_kernel" if args.precision == "bf16" else "epic_simt_kernel",
                             "kernel_ms": k_ms, "algorithmic_flop_per_launch": flops,
                             "peak_source": f"{peaks['source']} bf16_tflops_sustained (MEASURED_PEAKS.json)"}}
        if all_real is not None:
            line["roofline_all_real"] = {"frac": all_real["achieved"] / peak, "achieved": all_real["achieved"], "peak": peak,
                                         "unit": "TFLOP/s", "jets_per_s": all_real["jets_per_s"], "jets": all_real["jets"],
                                         "kernel_ms": all_real["kernel_ms"],
                                         "workload": "same kernel, every jet 150 real particles (SURVEY 8d all-real variant)"}
        if multi_check is not None:
            line["multi_gpu_check"] = multi_check
        if gpu_eager is not None:
            line["reference_gpu_eager"] = gpu_eager
        if train is not None:
            line["train"] = train
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            run = oracle_sampler()
            jps, secs, mean_n = time_cpu(run, args.cpu_baseline_jets, 9999)
            line["cpu_baseline"] = {"value": jps, "unit": "jets/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_baseline_jets} jets of the same workload (full 398 evaluations, "
                                              f"mean multiplicity {mean_n:.1f}), {secs:.1f} s, torch CPU fp32 oracle"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
