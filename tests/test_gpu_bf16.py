"""GPU: the bf16 tcgen05 path against the oracle.  Tolerances (north star): per-evaluation vector field
within 2e-2 relative (bf16 operands, fp32 accumulation / residual / state), end points within 1e-3."""
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import loss_oracle as lo
from oracle import ode_oracle as oo

from helpers import Golden, build_module, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16_STEP_TOL = 2e-2
BF16_END_TOL = 1e-3


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def make(ctor, sd):
    m = build_module(ctor, sd, device=DEV)
    m.set_precision("bf16")
    return m


def vf_cuda(m, t, x, cond, mask):
    with torch.no_grad():
        return m.flows[0](t.to(DEV), x.to(DEV), cond=None if cond is None else cond.to(DEV),
                          mask=None if mask is None else mask.to(DEV)).cpu()


@pytest.mark.parametrize("name", ["c1_jetnet30", "c2_jetnet150"])
def test_vector_field_bf16_vs_reference_golden(name):
    g = Golden(name)
    m = make(g.ctor, g.sd)
    N = g.x.shape[1]
    v_s = vf_cuda(m, g.t("t_sample"), g.x, None, g.mask)
    v_t = vf_cuda(m, g.t("t_train").unsqueeze(-1).repeat_interleave(N, dim=1), g.x, None, g.mask)
    es, et = rel_l2(v_s, g.t("v_sample")), rel_l2(v_t, g.t("v_train"))
    print(f"{name}: bf16 vector field rel-L2 sampling-mode {es:.2e}, training-mode {et:.2e}")
    assert es < BF16_STEP_TOL and et < BF16_STEP_TOL
    assert (v_s * (1 - g.mask)).abs().max() == 0


@pytest.mark.parametrize("name", ["c1_jetnet30", "c2_jetnet150"])
def test_sample_bf16_vs_reference_golden(name):
    g = Golden(name)
    m = make(g.ctor, g.sd)
    for solver, steps in g.meta["ode"]:
        torch.manual_seed(777)
        s = m.sample(g.x.shape[0], mask=g.mask, ode_solver=solver, ode_steps=steps).cpu()
        e = rel_l2(s, g.t(f"sample_{solver}{steps}"))
        print(f"{name}: {solver}-{steps} end point rel-L2 {e:.2e}")
        assert e < BF16_END_TOL
        assert (s * (1 - g.mask)).abs().max() == 0


def test_teacher_forced_per_step_bf16():
    """Every 5th evaluation of a Euler-100 JetNet-30 run on the oracle trajectory; gate with the norm floor of
    SURVEY B.14: |dv| <= tol * max(|v_ref|, median_k |v_ref,k|)."""
    g = Golden("c1_jetnet30")
    m = make(g.ctor, g.sd)
    z = g.t("z_euler100") * g.mask
    with torch.no_grad():
        _, evals = oo.integrate(g.oracle_vf(), z, 100, "euler", return_evals=True)
    norms = torch.tensor([float(v.double().norm()) for _, _, v in evals])
    floor = float(norms.median())
    worst_rel, worst_floor = 0.0, 0.0
    for t, x, v in evals[::5]:
        d = float((vf_cuda(m, t, x, None, g.mask).double() - v.double()).norm())
        worst_rel = max(worst_rel, d / float(v.double().norm()))
        worst_floor = max(worst_floor, d / max(float(v.double().norm()), floor))
    print(f"teacher-forced bf16: worst plain rel {worst_rel:.2e}, worst with norm floor {worst_floor:.2e}")
    assert worst_floor < BF16_STEP_TOL


def test_bf16_groups_edge_cases_and_conditioning():
    g = Golden("c2_jetnet150")
    # conditioned default-size net (fm_tops150_cond-like: global 2 / local 2), many groups, an empty jet
    ctor = {**g.ctor, "global_cond_dim": 2, "local_cond_dim": 2}
    cfg = eo.EpicCfg(**{**g.meta["cfg"], "global_cond_dim": 2, "local_cond_dim": 2})
    sd = eo.synth_state_dict(cfg, 77)
    m = make(ctor, sd)
    B = 97
    x, mask, cond = eo.synth_cloud(B, 150, 3, 4321, cond_dim=2)
    mask[5] = 0
    x = x * mask
    t = torch.tensor(0.421)
    v = vf_cuda(m, t, x, cond, mask)
    with torch.no_grad():
        vo = eo.cnf_forward(sd, cfg, t, x, cond, mask, **g.oracle_kwargs())
    keep = [i for i in range(B) if i != 5]
    assert torch.isnan(v[5]).all() and not torch.isnan(v[keep]).any()
    e = rel_l2(v[keep], vo[keep])
    print(f"conditioned bf16, {B} jets: rel-L2 {e:.2e}")
    assert e < BF16_STEP_TOL
    assert torch.equal(vf_cuda(m, t, x + (1 - mask) * 5.0, cond, mask)[keep], v[keep])
    # agreement with the fp32 CUDA path on a whole integration
    torch.manual_seed(3)
    a = m.sample(B, cond=cond, mask=mask, ode_solver="midpoint", ode_steps=10).cpu()
    m.set_precision("fp32")
    torch.manual_seed(3)
    b = m.sample(B, cond=cond, mask=mask, ode_solver="midpoint", ode_steps=10).cpu()
    assert rel_l2(a[keep], b[keep]) < BF16_END_TOL


def test_bf16_unsupported_shapes_fail_loudly():
    g = Golden("cond_lhco_like")          # hid 40
    m = build_module(g.ctor, g.sd, device=DEV)
    with pytest.raises(Exception, match="hid == 128"):
        m.set_precision("bf16")
        vf_cuda(m, g.t("t_sample"), g.x, g.cond, g.mask)
