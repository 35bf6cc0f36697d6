"""CPU, world_size 2 over gloo: host-side logic of the multi-GPU launchers.

The CUDA kernels cannot run here, so the per-rank integration is the CPU oracle (tests may use it) and the flat
gradient is a synthetic tensor: what is exercised is sharding, slicing of the shared noise, the final gather and
the flat-gradient averaging -- the parts that differ between 1 and N ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import Golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


class _Hp:
    num_particles = 30
    features = 3


class _FakeModel:
    hparams = _Hp()
    device = torch.device("cpu")


def _oracle_integrate(g, steps):
    from oracle import loss_oracle as lo

    def integrate(z, cond, mask):
        vf = g.oracle_vf(cond=cond, mask=mask)
        with torch.no_grad():
            return lo.sample(vf, z, None, "midpoint", steps)      # z is already masked by the launcher
    return integrate


def _gen_worker(rank, world, port, n, out_path):
    from particle_fm_b200.launch import generate_data_sharded
    _init(rank, world, port)
    torch.set_num_threads(2)
    g = Golden("c1_jetnet30")
    mask = torch.cat([g.mask] * 3)[:n]
    torch.manual_seed(2024)                                       # same CPU generator state on every rank
    res = generate_data_sharded(_FakeModel(), n, mask=mask, ode_solver="midpoint", ode_steps=4,
                                integrate=_oracle_integrate(g, 4))
    torch.manual_seed(2024)
    res_all = generate_data_sharded(_FakeModel(), n, mask=mask, ode_solver="midpoint", ode_steps=4, gather="all",
                                    integrate=_oracle_integrate(g, 4))
    assert res_all.shape == (n, 30, 3)
    if rank == 0:
        assert torch.equal(res, res_all)
        torch.save(res, out_path)
    else:
        assert res is None
    dist.destroy_process_group()


def test_sharded_generation_equals_single_process(tmp_path):
    from particle_fm_b200.launch import generate_data_sharded, shard_bounds
    n = 13                                                        # odd: the last rank is short
    g = Golden("c1_jetnet30")
    mask = torch.cat([g.mask] * 3)[:n]
    torch.manual_seed(2024)
    single = generate_data_sharded(_FakeModel(), n, mask=mask, ode_solver="midpoint", ode_steps=4,
                                   integrate=_oracle_integrate(g, 4))
    out = str(tmp_path / "res.pt")
    mp.spawn(_gen_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    multi = torch.load(out)
    assert multi.shape == single.shape == (n, 30, 3)
    # same noise slices, same jets: equal up to the CPU oracle's batch-size-dependent GEMM blocking
    assert torch.allclose(multi, single, rtol=1e-5, atol=1e-6)
    assert (multi * (1 - mask)).abs().max() == 0
    # bounds: contiguous cover, short / empty tail ranks
    for nn, w in [(13, 2), (5, 8), (16, 4), (1, 2)]:
        b = [shard_bounds(nn, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == nn and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _grad_worker(rank, world, port, out_path):
    from particle_fm_b200.launch import attach_flat_grad_allreduce, detach_flat_grad_allreduce
    _init(rank, world, port)

    class Net:
        flat_grad_hook = None

    class Flow:
        net = Net()

    class Model:
        flows = [Flow()]

    m = attach_flat_grad_allreduce(Model())
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)     # rank r holds (r+1) * [0..9]
    avg = m.flows[0].net.flat_grad_hook(flat)
    expect = torch.arange(10, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    assert torch.allclose(avg, expect)
    detach_flat_grad_allreduce(m)
    assert m.flows[0].net.flat_grad_hook is None
    if rank == 0:
        torch.save(avg, out_path)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_mean(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.allclose(torch.load(out), torch.arange(10, dtype=torch.float32) * 1.5)
