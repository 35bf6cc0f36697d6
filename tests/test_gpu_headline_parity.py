"""GPU: parity of the HEADLINE configuration at full length, and of the bf16 instantiations round 1 left untested.

(a) the FP=8 instantiation of the tcgen05 kernel (JetClass-uncond, 128 particles x 8 features) against the oracle;
(b) C2 (150 x 3, variable multiplicity) over the FULL midpoint ode_steps=200 integration bench.py times (398 evaluations):
    end point of the bf16 and fp32 paths against oracle.loss_oracle.sample, and the bf16 vector field teacher-forced on the
    oracle trajectory with the norm floor of SURVEY B.14;
(c) W1m / W1p on the 150-particle shape (north star: within the reference's seed-to-seed spread).
Tolerances (north star / SURVEY 8d): per-evaluation 2e-2 (bf16), end point 1e-3 (bf16) / 1e-4 (fp32)."""
import numpy as np
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import loss_oracle as lo
from oracle import metrics_oracle as mo
from oracle import ode_oracle as oo

from helpers import Golden, build_module, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16_STEP_TOL, BF16_END_TOL, FP32_END_TOL = 2e-2, 1e-3, 1e-4
KW = dict(t_emb="cosine", frequencies=16, add_time_to_input=False)


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def vf_cuda(m, t, x, cond, mask):
    with torch.no_grad():
        return m.flows[0](t.to(DEV), x.to(DEV), cond=None if cond is None else cond.to(DEV), mask=mask.to(DEV)).cpu()


def test_bf16_fp8_instantiation_jetclass_uncond():
    """epic_tc_kernel<8,...>: 8 per-particle features (BASELINE config 5, unconditional)."""
    N, F = 128, 8
    cfg = eo.EpicCfg(feats=F, input_dim=F, hid=128, latent=10, layers=6, t_dim=32, t_local_cat=True, t_global_cat=True)
    ctor = dict(features=F, hidden_dim=128, num_particles=N, frequencies=16, layers=6, latent=10, t_emb="cosine",
                t_local_cat=True, t_global_cat=True, add_time_to_input=False)
    sd = eo.synth_state_dict(cfg, 2024)
    m = build_module(ctor, sd, device=DEV).set_precision("bf16")
    B = 40
    x, mask, _ = eo.synth_cloud(B, N, F, 91)
    mask[3, 1:] = 0                                               # a one-particle jet
    mask[7] = 1                                                   # a full jet
    x = x * mask
    vf = lambda t, y: eo.cnf_forward(sd, cfg, t, y, None, mask, **KW)
    for tv in (0.93, 0.37, 0.02):
        t = torch.tensor(tv)
        with torch.no_grad():
            v_ref = vf(t, x)
        v = vf_cuda(m, t, x, None, mask)
        e = rel_l2(v, v_ref)
        print(f"FP=8 bf16 vector field t={tv}: rel-L2 {e:.2e}")
        assert e < BF16_STEP_TOL
        assert (v * (1 - mask)).abs().max() == 0
    with torch.no_grad():
        s_ref = lo.sample(vf, x, mask, "midpoint", 12)
        s = m.flows[0].decode(x.to(DEV), None, mask.to(DEV), "midpoint", 12).cpu()
    e = rel_l2(s, s_ref)
    print(f"FP=8 bf16 midpoint-12 end point rel-L2 {e:.2e}")
    assert e < BF16_END_TOL
    m.set_precision("fp32")
    with torch.no_grad():
        s32 = m.flows[0].decode(x.to(DEV), None, mask.to(DEV), "midpoint", 12).cpu()
    assert rel_l2(s32, s_ref) < FP32_END_TOL


@pytest.fixture(scope="module")
def headline():
    """32 jets of C2 integrated by the oracle over the full midpoint ode_steps=200 grid, every evaluation kept."""
    g = Golden("c2_jetnet150")
    B, N = 32, 150
    _, mask, _ = eo.synth_cloud(B, N, 3, 9999)
    mask[0] = 1                                                   # one full 150-particle jet (straddles two 128-row tiles)
    mask[1, 129:] = 0; mask[1, :129] = 1                          # 129 particles: one row into the second tile
    torch.manual_seed(4242)
    z = torch.randn(B, N, 3) * mask
    vf = g.oracle_vf(mask=mask)
    with torch.no_grad():
        end, evals = oo.integrate(vf, z, 200, "midpoint", return_evals=True)
    assert len(evals) == 398
    return g, mask, z, end, evals


def test_headline_full_length_end_point(headline):
    g, mask, z, end, _ = headline
    m = build_module(g.ctor, g.sd, device=DEV)
    for prec, tol in (("bf16", BF16_END_TOL), ("fp32", FP32_END_TOL)):
        m.set_precision(prec)
        with torch.no_grad():
            s = m.flows[0].decode(z.to(DEV), None, mask.to(DEV), "midpoint", 200).cpu()
        e = rel_l2(s, end)
        per_jet = ((s - end).double().flatten(1).norm(dim=1) / end.double().flatten(1).norm(dim=1)).max()
        print(f"C2 midpoint-200 (398 evaluations) end point, {prec}: rel-L2 {e:.2e}, worst jet {float(per_jet):.2e}")
        assert e < tol
        assert float(per_jet) < 5 * tol
        assert (s * (1 - mask)).abs().max() == 0
        assert torch.isfinite(s).all()


def test_headline_teacher_forced_per_step_bf16(headline):
    """Every 7th of the 398 evaluations on the oracle's own trajectory: |dv| <= 2e-2 * max(|v_ref|, median_k |v_ref,k|)."""
    g, mask, _, _, evals = headline
    m = build_module(g.ctor, g.sd, device=DEV).set_precision("bf16")
    norms = torch.tensor([float(v.double().norm()) for _, _, v in evals])
    floor = float(norms.median())
    worst_rel, worst_floor = 0.0, 0.0
    for t, x, v in evals[::7]:
        d = float((vf_cuda(m, t, x, None, mask).double() - v.double()).norm())
        worst_rel = max(worst_rel, d / float(v.double().norm()))
        worst_floor = max(worst_floor, d / max(float(v.double().norm()), floor))
    print(f"C2 teacher-forced bf16 over 398 evaluations: worst plain rel {worst_rel:.2e}, with norm floor {worst_floor:.2e}")
    assert worst_floor < BF16_STEP_TOL


def test_w1_metrics_jetnet150_shape():
    """W1m / W1p of CUDA samples (same noise as oracle run A) against the spread between oracle runs with different
    noise seeds, on the 150-particle shape the north star names."""
    g = Golden("c2_jetnet150")
    B, N, steps = 320, 150, 6
    _, mask, _ = eo.synth_cloud(B, N, 3, 2718)
    vf = g.oracle_vf(mask=mask)

    def oracle_run(seed):
        torch.manual_seed(seed)
        z = torch.randn(B, N, 3)
        with torch.no_grad():
            return z, lo.sample(vf, z, mask, "midpoint", steps).numpy()

    zA, oA = oracle_run(1)
    mk = mask.squeeze(-1).numpy()
    sm, sp = [], []
    for seed in (2, 3, 4, 5, 6):
        _, oS = oracle_run(seed)
        sm.append(mo.w1m(oA, oS)[0])
        sp.append(mo.w1p(oA, mk, oS, mk)[0])
    spread_m, spread_p = float(np.mean(sm)), float(np.mean(sp))
    m = build_module(g.ctor, g.sd, device=DEV)
    for prec in ("fp32", "bf16"):
        m.set_precision(prec)
        with torch.no_grad():
            s = m.flows[0].decode((zA * mask).to(DEV), None, mask.to(DEV), "midpoint", steps).cpu().numpy()
        wm, wp = mo.w1m(oA, s)[0], mo.w1p(oA, mk, s, mk)[0]
        floor_m, floor_p = mo.w1m(oA, oA)[0], mo.w1p(oA, mk, oA, mk)[0]
        print(f"N=150 {prec}: W1m {wm:.3e} (floor {floor_m:.3e}, seed-to-seed {spread_m:.3e}); "
              f"W1p {wp:.3e} (floor {floor_p:.3e}, seed-to-seed {spread_p:.3e})")
        assert wm <= max(spread_m, 1.5 * floor_m) and wp <= max(spread_p, 1.5 * floor_p)
