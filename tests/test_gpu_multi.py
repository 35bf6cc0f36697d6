"""GPU, world_size 2 over NCCL (skipped on a one-GPU box): the product launcher particle_fm_b200.launch.generate_data_sharded
with the real fused kernels.  Rank 0 checks the gathered result against its own single-GPU re-integration of every rank's
noise slice (bit-equal: same inputs, same plan, deterministic kernels) and against one single-GPU call over the whole
request (the plan differs, so equal to fp32 rounding), in both noise modes."""
import os
import socket

import pytest
import torch

from helpers import Golden, build_module

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, precision, noise):
    import torch.distributed as dist
    from particle_fm_b200.launch import block_noise, generate_data_sharded, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = Golden("c2_jetnet150")
    m = build_module(g.ctor, g.sd, device=dev).set_precision(precision)
    n, N, F, steps = 301, 150, 3, 6                                  # odd: the last rank is short
    mask = torch.cat([g.mask] * (n // g.mask.shape[0] + 1))[:n]
    torch.manual_seed(2024)
    res = generate_data_sharded(m, n, mask=mask, ode_solver="midpoint", ode_steps=steps, noise=noise)
    if rank == 0:
        assert res.shape == (n, N, F)
        torch.manual_seed(2024)
        if noise == "stream":
            z = torch.randn(n, N, F)
        else:
            z = block_noise(0, n, N, F, int(torch.randint(0, 2 ** 62, (1,))))
        z = z * mask
        with torch.no_grad():
            for r in range(world):                                   # every rank's slice, re-integrated on this GPU
                lo, hi = shard_bounds(n, world, r)
                ref = m.forward(z[lo:hi].to(dev), mask=mask[lo:hi].to(dev), reverse=True, ode_solver="midpoint",
                                ode_steps=steps).cpu()
                assert torch.equal(res[lo:hi], ref), f"rank {r} slice differs from its single-GPU re-integration"
            whole = m.forward(z.to(dev), mask=mask.to(dev), reverse=True, ode_solver="midpoint", ode_steps=steps).cpu()
        err = float((res - whole).abs().max())
        print(f"{precision}/{noise}: sharded vs one single-GPU call over the whole request, max |d| = {err:.2e}")
        assert err < (2e-3 if precision == "bf16" else 1e-5)
        assert (res * (1 - mask)).abs().max() == 0
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision,noise", [("fp32", "stream"), ("bf16", "blocks")])
def test_sharded_generation_on_two_gpus(lib_built, precision, noise):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), precision, noise), nprocs=2, join=True)
