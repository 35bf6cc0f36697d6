"""GPU: the CUDA path (through the C ABI) against the oracle and the golden vectors recorded from the
reference.  fp32 path tolerance: relative L2 <= 1e-5 per evaluation (north star: 1e-3), end points of
whole integrations <= 1e-4 relative L2."""
import numpy as np
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import loss_oracle as lo
from oracle import ode_oracle as oo

from helpers import GOLDEN_CASES, Golden, build_module, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_STEP_TOL = 1e-5
FP32_END_TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def cuda_vf(m):
    cnf = m.flows[0]
    def f(t, x, cond=None, mask=None):
        with torch.no_grad():
            return cnf(t.to(DEV), x.to(DEV), cond=None if cond is None else cond.to(DEV),
                       mask=None if mask is None else mask.to(DEV)).cpu()
    return f


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_vector_field_vs_reference_golden(name):
    g = Golden(name)
    m = build_module(g.ctor, g.sd, device=DEV)
    f = cuda_vf(m)
    N = g.x.shape[1]
    v_s = f(g.t("t_sample"), g.x, g.cond, g.mask)
    tbn = g.t("t_train").unsqueeze(-1).repeat_interleave(N, dim=1)
    v_t = f(tbn, g.x, g.cond, g.mask)
    assert rel_l2(v_s, g.t("v_sample")) < FP32_STEP_TOL
    assert rel_l2(v_t, g.t("v_train")) < FP32_STEP_TOL
    assert (v_s * (1 - g.mask)).abs().max() == 0
    # int64 mask as the datamodules deliver it (jetnet_datamodule.py:247-249)
    assert torch.equal(f(g.t("t_sample"), g.x, g.cond, g.mask.long()), v_s)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sample_vs_reference_golden(name):
    """Whole integrations: same CPU-generator noise as the reference run that produced the golden."""
    g = Golden(name)
    m = build_module(g.ctor, g.sd, device=DEV)
    B = g.x.shape[0]
    for solver, steps in g.meta["ode"]:
        torch.manual_seed(777)
        s = m.sample(B, cond=g.cond, mask=g.mask, ode_solver=solver, ode_steps=steps).cpu()
        ref = g.t(f"sample_{solver}{steps}")
        assert s.shape == ref.shape
        assert rel_l2(s, ref) < FP32_END_TOL, (solver, steps, rel_l2(s, ref))
        assert (s * (1 - g.mask)).abs().max() == 0
        assert m.flows[0].net.engine().last_launches() <= 5      # time/cond bias tables + 2 plan kernels + ONE integration kernel


def test_teacher_forced_per_step_parity():
    """Every network evaluation of a Euler-100 JetNet-30 run, fed the oracle's own (t_k, x_k)."""
    g = Golden("c1_jetnet30")
    m = build_module(g.ctor, g.sd, device=DEV)
    f = cuda_vf(m)
    z = g.t("z_euler100") * g.mask
    with torch.no_grad():
        _, evals = oo.integrate(g.oracle_vf(), z, 100, "euler", return_evals=True)
    worst = 0.0
    for t, x, v in evals[::7] + evals[-3:]:
        worst = max(worst, rel_l2(f(t, x, None, g.mask), v))
    assert worst < FP32_STEP_TOL, worst


def test_edge_cases():
    g = Golden("c1_jetnet30")
    m = build_module(g.ctor, g.sd, device=DEV)
    f = cuda_vf(m)
    t = torch.tensor(0.25)
    # zero-multiplicity jet -> NaN confined to that jet (epic.py:161,370)
    x, mask = g.x.clone(), g.mask.clone()
    mask[1] = 0
    x = x * mask
    v = f(t, x, None, mask)
    vo = g.oracle_vf(mask=mask)(t, x)
    assert torch.isnan(v[1]).all() and not torch.isnan(v[[0, 2, 3, 4, 5]]).any()
    keep = [0, 2, 3, 4, 5]
    assert rel_l2(v[keep], vo[keep]) < FP32_STEP_TOL
    # mask=None (fixed-size sets), single jet, single particle
    v = f(t, g.x, None, None)
    assert rel_l2(v, g.oracle_vf(mask=None)(t, g.x)) < FP32_STEP_TOL
    v1 = f(t, g.x[:1], None, g.mask[:1])
    assert rel_l2(v1, eo.cnf_forward(g.sd, g.cfg, t, g.x[:1], None, g.mask[:1], **g.oracle_kwargs())) < FP32_STEP_TOL
    one = torch.zeros(2, 30, 1)
    one[:, 0] = 1
    xo = g.x[:2] * one
    assert rel_l2(f(t, xo, None, one), eo.cnf_forward(g.sd, g.cfg, t, xo, None, one, **g.oracle_kwargs())) < FP32_STEP_TOL
    # ragged (non-prefix) masks and garbage in the padded slots
    x, mask, _ = eo.synth_cloud(9, 30, 3, 5, ragged=True)
    junk = x + (1 - mask) * 3.3
    vo = eo.cnf_forward(g.sd, g.cfg, t, x, None, mask, **g.oracle_kwargs())
    assert rel_l2(f(t, junk, None, mask), vo) < FP32_STEP_TOL


@pytest.mark.parametrize("B,N", [(257, 30), (64, 150), (40, 279)])
def test_many_groups_and_properties(B, N):
    """More jets than fit one CTA: grouping must not mix jets.  Oracle on a subset + properties on all:
    padding independence, permutation equivariance, batch-order independence."""
    g = Golden("c2_jetnet150")
    cfg = eo.EpicCfg(**{**g.meta["cfg"]})
    ctor = {**g.ctor, "num_particles": N}
    m = build_module(ctor, g.sd, device=DEV)
    f = cuda_vf(m)
    x, mask, _ = eo.synth_cloud(B, N, 3, 900 + N)
    t = torch.tensor(0.613)
    v = f(t, x, None, mask)
    sub = torch.arange(0, B, max(1, B // 8))
    vo = eo.cnf_forward(g.sd, cfg, t, x[sub], None, mask[sub], **g.oracle_kwargs())
    assert rel_l2(v[sub], vo) < FP32_STEP_TOL
    assert torch.equal(f(t, x + (1 - mask) * 9.0, None, mask), v)
    pb = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    vb = f(t, x[pb], None, mask[pb])
    assert rel_l2(vb, v[pb]) < 2e-6          # different grouping -> same jets, fp32 re-association only
    pn = torch.randperm(N, generator=torch.Generator().manual_seed(2))
    vn = f(t, x[:, pn], None, mask[:, pn])
    assert rel_l2(vn, v[:, pn]) < 2e-6


def test_full_size_shapes_run_and_stay_finite():
    """BASELINE shapes that the oracle cannot finish quickly: JetNet-150 batch 1000 midpoint (short grid),
    checked through invariants (finite, padded zeros, deterministic, independent of batch split)."""
    g = Golden("c2_jetnet150")
    m = build_module(g.ctor, g.sd, device=DEV)
    B = 1000
    _, mask, _ = eo.synth_cloud(B, 150, 3, 9999)
    torch.manual_seed(5)
    a = m.sample(B, mask=mask, ode_solver="midpoint", ode_steps=6).cpu()
    torch.manual_seed(5)
    z = torch.randn(B, 150, 3)
    lo_half = m.flows[0].decode((z * mask)[:500].to(DEV), None, mask[:500].to(DEV), "midpoint", 6).cpu()
    assert torch.isfinite(a).all() and (a * (1 - mask)).abs().max() == 0
    assert rel_l2(a[:500], lo_half) < 2e-6
    with torch.no_grad():
        ref = lo.sample(g.oracle_vf(mask=mask[:6]), z[:6], mask[:6], "midpoint", 6)
    assert rel_l2(a[:6], ref) < FP32_END_TOL


def test_weights_refresh_after_load_state_dict():
    """EMA swaps weights with load_state_dict between calls (ema.py:145-159): the packed copy must follow."""
    g = Golden("tglobal_plain")
    m = build_module(g.ctor, g.sd, device=DEV)
    f = cuda_vf(m)
    t = g.t("t_sample")
    v0 = f(t, g.x, None, g.mask)
    sd2 = eo.synth_state_dict(g.cfg, 4321, weight_norm=False)
    from helpers import full_state_dict
    m.load_state_dict(full_state_dict(m, sd2))
    v1 = f(t, g.x, None, g.mask)
    assert rel_l2(v1, eo.cnf_forward(sd2, g.cfg, t, g.x, None, g.mask, **g.oracle_kwargs())) < FP32_STEP_TOL
    assert rel_l2(v1, v0) > 1e-2


def test_generate_data_end_to_end():
    from particle_fm_b200.utils.data_generation import generate_data
    g = Golden("c1_jetnet30")
    m = build_module(g.ctor, g.sd, device=DEV)
    n = 70
    _, mask, _ = eo.synth_cloud(n, 30, 3, 31)
    torch.manual_seed(99)
    out, secs = generate_data(m, n, batch_size=32, device=DEV, variable_set_sizes=True, mask=mask, normalized_data=True,
                              means=[0.1, 0.2, 0.3], stds=[1.0, 2.0, 3.0], verbose=False, ode_solver="euler", ode_steps=9)
    assert out.shape == (n, 30, 3) and secs >= 0
    torch.manual_seed(99)
    zs = [torch.randn(32, 30, 3), torch.randn(32, 30, 3), torch.randn(6, 30, 3)]
    z = torch.cat(zs)
    mk = torch.cat([mask[:32], mask[32:64], mask[-6:]])
    with torch.no_grad():
        ref = lo.sample(g.oracle_vf(mask=mk), z, mk, "euler", 9)
    ref = ref * torch.tensor([1.0, 2.0, 3.0]) / 5 + torch.tensor([0.1, 0.2, 0.3])
    ref = ref * mk
    assert rel_l2(torch.from_numpy(out), ref) < FP32_END_TOL
