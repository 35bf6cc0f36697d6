"""CPU: host-side mirror of the reference interface (state_dict layout, init stream, time grids,
unsupported-option errors, generate_data batching)."""
import numpy as np
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import ode_oracle as oo
from oracle import ref_shim

from helpers import GOLDEN_CASES, Golden, build_module

YAML_DEFAULT = dict(features=3, hidden_dim=128, num_particles=30, frequencies=16, layers=6, latent=10,
                    t_emb="cosine", t_local_cat=True, t_global_cat=True, add_time_to_input=False)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_state_dict_layout_matches_reference(name):
    """Keys, order and shapes: flows.0.net.<linear>.{bias,weight_g,weight_v}, flows.0.frequencies and
    the duplicate loss.flows.0.* entries the reference checkpoints carry."""
    g = Golden(name)
    m = build_module(g.ctor, g.sd)
    keys = list(m.state_dict().keys())
    names = [str(n) for n in g.arr["grad_fmot_names"]]           # reference named_parameters() order
    expect = ["flows.0.frequencies"] + ["flows.0.net." + n for n in names]
    expect += ["loss." + k for k in expect]
    assert keys == expect
    shapes = eo.linear_shapes(g.cfg)
    sd = m.state_dict()
    for n in names:
        lin, kind = n.rsplit(".", 1)
        o, i = shapes[lin]
        want = {"bias": (o,), "weight_g": (o, 1), "weight_v": (o, i), "weight": (o, i)}[kind]
        assert tuple(sd["flows.0.net." + n].shape) == want
    assert [k for k, _ in m.flows[0].net.named_parameters()] == names


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_default_init_consumes_rng_like_reference():
    ref = ref_shim.load()
    from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
    torch.manual_seed(12345)
    a = ref.fm.SetFlowMatchingLitModule(optimizer=None, **YAML_DEFAULT)
    torch.manual_seed(12345)
    b = SetFlowMatchingLitModule(optimizer=None, **YAML_DEFAULT)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    b.load_state_dict(sa, strict=True)        # reference checkpoints load unchanged


@pytest.mark.parametrize("solver", ["euler", "midpoint"])
@pytest.mark.parametrize("steps", [2, 3, 100, 200])
def test_fixed_step_grid_equals_oracle(solver, steps):
    from particle_fm_b200.models.flow_matching_module import fixed_step_grid
    t, dt = fixed_step_grid(steps, solver)
    to, dto = oo.time_grid(steps, solver)
    assert torch.equal(t, to) and torch.equal(dt, dto)
    assert t.dtype == torch.float32


def test_time_codes_equal_oracle():
    from particle_fm_b200.models.flow_matching_module import CNF
    cnf = CNF(frequencies=16, hidden_dim=8, layers=1, latent=4, t_emb="cosine", t_local_cat=True, t_global_cat=True,
              add_time_to_input=False)
    t, _ = oo.time_grid(200, "midpoint")
    assert torch.equal(cnf.time_code(t), eo.cosine_time_code(t, 32))
    for ti in t[:5]:     # batched table == the per-call 0-dim evaluation the reference performs
        assert torch.equal(cnf.time_code(ti)[0], eo.cosine_time_code(ti.unsqueeze(0), 32)[0])
    cnf2 = CNF(frequencies=6, hidden_dim=8, layers=1, latent=4)
    assert torch.equal(cnf2.time_code(t), eo.sincos_time_code(t, 6))


def test_unsupported_options_raise():
    from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule as M
    small = dict(hidden_dim=8, layers=1, latent=4, frequencies=2)
    for kw in (dict(model="mdma"), dict(model="droid_fulltransformer"), dict(use_normaliser=True),
               dict(t_emb="gaussian"), dict(dropout=0.1), dict(loss_type="no_such_loss"),
               dict(n_transforms=2), dict(activation="relu"), dict(wrapper_func="spectral_norm")):
        with pytest.raises(NotImplementedError):
            M(optimizer=None, **{**small, **kw})
    m = M(optimizer=None, **small)
    for solver in ("dopri5_zuko", "rk4", "dopri5", "tsit5"):
        with pytest.raises(NotImplementedError):
            m.forward(torch.zeros(1, 3, 3), reverse=True, ode_solver=solver)
    with pytest.raises(SyntaxError):       # "em" / "ddim" without the diffusion loss: flow_matching_module.py:328-329
        m.forward(torch.zeros(1, 3, 3), reverse=True, ode_solver="em")
    assert m.hparams.num_particles == 150 and m.hparams.loss_type == "FM-OT" and m.conditioned is False


def test_generate_data_batching_and_postprocessing():
    """Same batching / remainder / mask / inverse-normalisation logic as data_generation.py:77-172."""
    from particle_fm_b200.utils.data_generation import generate_data

    class Fake(torch.nn.Module):
        calls = []

        def sample(self, n_samples, cond=None, mask=None, ode_solver="midpoint", ode_steps=100):
            Fake.calls.append((n_samples, None if cond is None else cond.clone(), ode_solver, ode_steps))
            return torch.ones(n_samples, 4, 3)

    mask = (torch.arange(4)[None, :, None] < torch.tensor([1, 2, 3, 4, 1, 2, 3])[:, None, None]).float()
    cond = torch.arange(7.0)[:, None]
    out, dt = generate_data(Fake(), 7, batch_size=3, cond=cond, device="cpu", variable_set_sizes=True, mask=mask,
                            normalized_data=True, normalize_sigma=5, means=[1.0, 2.0, 3.0], stds=[5.0, 10.0, 15.0],
                            verbose=False, ode_solver="euler", ode_steps=7)
    assert out.shape == (7, 4, 3) and dt >= 0
    assert [c[0] for c in Fake.calls] == [3, 3, 1]
    assert torch.equal(Fake.calls[2][1], cond[-1:]) and Fake.calls[0][2:] == ("euler", 7)
    want = np.array([2.0, 4.0, 6.0])[None, None, :] * mask.numpy()
    assert np.allclose(out, want)
    with pytest.raises(ValueError):
        generate_data(Fake(), 7, variable_set_sizes=True, mask=None, device="cpu")
    with pytest.raises(ValueError):
        generate_data(Fake(), 6, mask=mask, device="cpu")


def test_block_noise_is_independent_of_the_partition():
    """launch.block_noise: the rows a rank draws for its slice equal the same rows of the whole request, whatever the
    world size (fixed-size blocks with their own seeds), so noise="blocks" results do not depend on the GPU count."""
    from particle_fm_b200.launch import block_noise, shard_bounds
    from particle_fm_b200.launch import generate as gen
    old = gen.NOISE_BLOCK
    gen.NOISE_BLOCK = 8
    try:
        n, N, F, seed = 37, 5, 3, 123456789
        whole = block_noise(0, n, N, F, seed)
        assert whole.shape == (n, N, F) and float(whole.std()) > 0.5
        for world in (1, 2, 3, 8):
            parts = [block_noise(*shard_bounds(n, world, r), N, F, seed) for r in range(world)]
            assert torch.equal(torch.cat(parts), whole)
        assert not torch.equal(block_noise(0, n, N, F, seed + 1), whole)
    finally:
        gen.NOISE_BLOCK = old


def test_training_fast_path_pieces_fail_loudly_without_cuda():
    """The fused optimizer and the graphed step are CUDA-only (no CPU fallback); the direct-gradient switch nests and restores."""
    from particle_fm_b200 import _lib, training
    from particle_fm_b200.launch import GraphedTrainStep
    from particle_fm_b200.optim import FusedClipAdamW
    w = torch.nn.Parameter(torch.ones(4))
    opt = FusedClipAdamW([w], lr=1e-3)
    w.grad = torch.ones(4)
    with pytest.raises(_lib.PfmError):
        opt.step()
    opt.zero_grad(set_to_none=True)
    assert opt.step() is None                           # nothing to do without gradients
    with pytest.raises(NotImplementedError):
        FusedClipAdamW([{"params": [w]}, {"params": [torch.nn.Parameter(torch.ones(2))]}])
    with pytest.raises(RuntimeError):
        GraphedTrainStep(None, opt, torch.zeros(2, 3, 3), torch.ones(2, 3, 1))
    assert training._DIRECT_GRADS is False
    with training.direct_param_grads():
        assert training._DIRECT_GRADS is True
        with training.direct_param_grads():
            assert training._DIRECT_GRADS is True
        assert training._DIRECT_GRADS is True
    assert training._DIRECT_GRADS is False
