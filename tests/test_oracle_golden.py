"""CPU: the oracle (plain-torch restatement) against the golden vectors recorded from the
UNMODIFIED reference (oracle/make_golden.py), and against the reference itself when the tree is
present (build container only)."""
import numpy as np
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import loss_oracle as lo
from oracle import ode_oracle as oo
from oracle import ref_shim

from helpers import GOLDEN_CASES, Golden


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)      # goldens were recorded single-threaded (reduction order)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_vector_field_matches_reference_golden(name):
    g = Golden(name)
    vf = g.oracle_vf()
    N = g.x.shape[1]
    with torch.no_grad():
        v_s = vf(g.t("t_sample"), g.x)
        v_t = vf(g.t("t_train").unsqueeze(-1).repeat_interleave(N, dim=1), g.x)
    # bit-exact on the recording platform; tolerate libm/BLAS differences elsewhere
    assert torch.allclose(v_s, g.t("v_sample"), rtol=1e-5, atol=1e-6)
    assert torch.allclose(v_t, g.t("v_train"), rtol=1e-5, atol=1e-6)
    assert (v_s * (1 - g.mask)).abs().max() == 0      # padded outputs are exactly zero (epic.py:391)


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("kind", ["FM-OT", "CFM", "droid"])
def test_losses_match_reference_golden(name, kind):
    g = Golden(name)
    tag = kind.replace("-", "").lower()
    n1 = g.t(f"loss_{tag}_n1") if f"loss_{tag}_n1" in g.arr else None
    sd = {k: v.clone().requires_grad_(True) for k, v in g.sd.items()}
    loss = lo.fm_loss(g.oracle_vf(sd=sd), kind, g.x, g.mask, g.t(f"loss_{tag}_t"), g.t(f"loss_{tag}_n0"), n1,
                      g.meta["sigma"])
    assert torch.allclose(loss.detach(), g.t(f"loss_{tag}"), rtol=1e-5)
    if kind == "FM-OT":
        loss.backward()
        names = [str(n) for n in g.arr["grad_fmot_names"]]
        norms = np.array([float(sd[n].grad.norm()) for n in names])
        assert np.allclose(norms, g.arr["grad_fmot_norms"], rtol=1e-4, atol=1e-7)
        if "grad_fmot" in g.arr:
            flat = torch.cat([sd[n].grad.flatten() for n in names])
            assert torch.allclose(flat, g.t("grad_fmot"), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sample_matches_reference_golden(name):
    g = Golden(name)
    vf = g.oracle_vf()
    for solver, steps in g.meta["ode"]:
        with torch.no_grad():
            s = lo.sample(vf, g.t(f"z_{solver}{steps}"), g.mask, solver, steps)
        ref = g.t(f"sample_{solver}{steps}")
        assert torch.allclose(s, ref, rtol=1e-4, atol=1e-5), (solver, steps, (s - ref).abs().max())


def test_loss_random_draw_order():
    """t from the CPU generator first, then the noise draws (losses.py:46-53, :104-116)."""
    x = torch.zeros(5, 7, 3)
    torch.manual_seed(3)
    t, n0, n1 = lo.draw_loss_randoms("CFM", x)
    torch.manual_seed(3)
    t2 = torch.rand(5)
    a = torch.randn(5, 7, 3)
    b = torch.randn(5, 7, 3)
    assert torch.equal(t, t2) and torch.equal(n0, a) and torch.equal(n1, b)


def test_time_grid_semantics():
    """ode_steps counts grid points; midpoint evaluates at t and t + dt/2; times run 1 -> 0."""
    for solver, per in (("euler", 1), ("midpoint", 2)):
        t, dt = oo.time_grid(100, solver)
        assert t.numel() == 99 * per and dt.numel() == 99
        assert t[0] == 1.0 and t.min() > 0 and torch.all(dt > 0)
        assert abs(float(dt.double().sum()) - 1.0) < 1e-5
    # closed-form grid == loop form used by the torchdyn stand-in (independent transcription)
    seen = []
    oo.integrate(lambda t, x: (seen.append(float(t)), torch.zeros_like(x))[1], torch.zeros(1, 2, 3), 12, "midpoint")
    assert np.allclose(seen, oo.time_grid(12, "midpoint")[0].numpy(), rtol=0, atol=0)


def test_zero_multiplicity_jet_is_nan_only_there():
    g = Golden("c1_jetnet30")
    x, mask = g.x.clone(), g.mask.clone()
    mask[1] = 0
    x = x * mask
    v = g.oracle_vf(mask=mask)(torch.tensor(0.25), x)
    assert torch.isnan(v[1]).all() and not torch.isnan(v[[0, 2, 3]]).any()


def test_padding_independence_and_equivariance():
    g = Golden("cond_lhco_like")
    vf = g.oracle_vf()
    t = torch.tensor(0.4)
    v = vf(t, g.x)
    junk = g.x + (1 - g.mask) * 7.7
    assert torch.equal(vf(t, junk), v)                 # padded inputs never matter
    perm = torch.randperm(g.x.shape[1], generator=torch.Generator().manual_seed(0))
    vp = eo.cnf_forward(g.sd, g.cfg, t, g.x[:, perm], g.cond, g.mask[:, perm], **g.oracle_kwargs())
    assert torch.allclose(vp, v[:, perm], atol=2e-6)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name", ["c1_jetnet30", "cond_lhco_like", "bare_sincos"])
def test_oracle_equals_live_reference(name):
    """Build container only: run the unmodified reference module and compare bit-for-bit."""
    from oracle.make_golden import build_reference
    ref = ref_shim.load()
    g = Golden(name)
    m = build_reference(ref, g.ctor, g.sd, "FM-OT")
    t = torch.tensor(0.3117)
    with torch.no_grad():
        v_ref = m.flows[0](t, g.x, cond=g.cond, mask=g.mask)
        assert torch.equal(v_ref, g.oracle_vf()(t, g.x))
        torch.manual_seed(11)
        s_ref = m.sample(g.x.shape[0], cond=g.cond, mask=g.mask, ode_solver="midpoint", ode_steps=5)
        torch.manual_seed(11)
        z = torch.randn_like(g.x)
        assert torch.equal(s_ref, lo.sample(g.oracle_vf(), z, g.mask, "midpoint", 5))
