"""GPU: the scope table's "next" rows through the C ABI against the reference-pinned goldens / the oracle.

  f2  pfm_postprocess + generate_data writing into one pinned host buffer      (data_generation.py:105-123)
  f1  pfm_ot_assign / pfm_ot_gather + ConditionalFlowMatchingOTLoss            (losses.py:140-204)
  f3  pfm_mlp_forward / pfm_mlp_sample + FLowMatchingNoSetsLitModule            (flow_matching_no_sets.py, mlp.py)
  f4  DiffusionLoss, pfm_epic_sample_diffusion (ddim / em / probability-flow)   (losses.py:207-285, solver.py)
Tolerances: integer / index work bit-exact; affine post-processing bit-exact, exp() column 2 ulp; fp32 network
results 1e-5 per evaluation, 1e-4 end points and gradients (the bar of the existing fp32 tests)."""
import json
import os
import types

import numpy as np
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import next_oracle as no

from helpers import GOLDEN_DIR, build_module, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files if k != "meta"}, json.loads(str(z["meta"]))


# ------------------------------------------------------------------------------------------------ f2
class _StubModel:
    """sample() hands back recorded raw samples on the device, like SetFlowMatchingLitModule.sample does."""

    def __init__(self, raw, N, F):
        self.raw, self.pos = raw.to(DEV), 0
        self.hparams = types.SimpleNamespace(num_particles=N, features=F)

    def to(self, device):
        return self

    def sample(self, n_samples, cond=None, mask=None, ode_solver="midpoint", ode_steps=100):
        out = self.raw[self.pos:self.pos + n_samples].clone()
        self.pos += n_samples
        return out


def test_generate_data_device_postprocessing_vs_reference_golden():
    from particle_fm_b200.utils.data_generation import generate_data
    arr, meta = _load("next_post")
    for name, c in meta.items():
        kw = c["kw"]
        raw, mask = torch.from_numpy(arr[f"{name}_raw"]), torch.from_numpy(arr[f"{name}_mask"])
        use_mask = kw.get("variable_set_sizes", False)
        data, secs = generate_data(_StubModel(raw, c["N"], c["F"]), c["n"], batch_size=c["batch"], device=DEV,
                                   mask=mask if use_mask else None, means=arr[f"{name}_means"], stds=arr[f"{name}_stds"],
                                   verbose=False, **kw)
        want = arr[f"{name}_out"]
        assert data.shape == want.shape and data.dtype == np.float32
        if kw.get("log_pt", False):
            cols = [f for f in range(c["F"]) if f != 2]
            assert np.array_equal(data[..., cols], want[..., cols]), name
            np.testing.assert_allclose(data[..., 2], want[..., 2], rtol=3e-6, atol=1e-6)      # expf vs numpy's exp
        else:
            assert np.array_equal(data, want), name                                          # bit-equal (signed zeros compare equal)


def test_postprocess_writes_pinned_host_memory_directly():
    from particle_fm_b200.engine import postprocess_into
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 150, 3, generator=g)
    mask = (torch.rand(1000, 150, 1, generator=g) > 0.4).float()
    scale, shift = [0.3, 1.7, 0.01], [0.5, -2.0, 3.0]
    out = torch.empty(1000, 150, 3, pin_memory=True)
    out.fill_(float("nan"))
    postprocess_into(x.to(DEV), mask.to(DEV), out, scale, shift, -1)
    torch.cuda.synchronize()
    want = (x * torch.tensor(scale) + torch.tensor(shift)) * mask
    assert torch.equal(out, want)
    with pytest.raises(ValueError, match="pinned"):
        postprocess_into(x.to(DEV), None, torch.empty(1000, 150, 3), scale, shift, -1)


# ------------------------------------------------------------------------------------------------ f1
@pytest.mark.parametrize("B,N,F,pad", [(7, 30, 3, True), (5, 150, 3, True), (3, 279, 3, False), (4, 33, 8, False), (2, 1, 3, False)])
def test_ot_assign_is_the_exact_assignment(B, N, F, pad):
    from scipy.optimize import linear_sum_assignment
    from particle_fm_b200.engine import ot_assign
    rs = np.random.RandomState(B * 1000 + N)
    x0 = rs.normal(size=(B, N, F)).astype("float32")
    x1 = (rs.normal(size=(B, N, F)) * 2).astype("float32")
    if pad:                                               # zero padding takes part in the transport like in the reference
        for k in range(B):
            x1[k, rs.randint(1, N + 1):] = 0
    sigma, cost = ot_assign(torch.from_numpy(x0).to(DEV), torch.from_numpy(x1).to(DEV), want_cost=True)
    sigma, cost = sigma.cpu().numpy(), cost.cpu().numpy()
    for k in range(B):
        M = ((x0[k].astype("float64")[:, None, :] - x1[k].astype("float64")[None, :, :]) ** 2).sum(-1)
        r, c = linear_sum_assignment(M)
        assert sorted(sigma[k]) == list(range(N))                           # a permutation
        assert abs(cost[k] - M[r, c].sum()) <= 1e-10 * max(1.0, M[r, c].sum()), (k, cost[k], M[r, c].sum())
        assert abs(M[np.arange(N), sigma[k]].sum() - M[r, c].sum()) <= 1e-10 * max(1.0, M[r, c].sum())
        if not pad:
            assert np.array_equal(sigma[k], c)                              # generic data: the optimum is unique


def test_cfm_ot_loss_vs_golden():
    arr, meta = _load("next_cfmot")
    for name, c in meta.items():
        cfg = eo.EpicCfg(**c["cfg"])
        sd = eo.synth_state_dict(cfg, c["wseed"])
        m = build_module(c["ctor"], sd, loss_type="CFM-OT", sigma=c["sigma"], device=DEV)
        g = lambda k: torch.from_numpy(arr[f"{name}_{k}"])
        x = g("x").to(DEV)
        draws = (g("x0").to(DEV), g("t"), arr[f"{name}_u"], g("eps").to(DEV))
        loss = m.loss(x, mask=g("mask").to(DEV), cond=None, draws=draws)
        loss.backward()
        assert np.array_equal(x.cpu().numpy(), arr[f"{name}_x1_after"]), name          # in-place re-indexing, bit-exact gather
        ref = float(arr[f"{name}_loss"])
        assert abs(float(loss) - ref) <= 1e-5 * abs(ref), (name, float(loss), ref)
        names = [str(s) for s in arr[f"{name}_grad_names"]]
        grads = {k[len("flows.0.net."):]: p.grad for k, p in m.named_parameters() if k.startswith("flows.0.net.")}
        for k, want in zip(names, arr[f"{name}_grad_norms"]):
            got = float(grads[k].norm())
            assert abs(got - want) <= 1e-4 * max(want, 1e-6), (name, k, got, want)


def test_cfm_ot_training_step_runs_with_its_own_draws():
    arr, meta = _load("next_cfmot")
    c = meta["ot_small"]
    cfg = eo.EpicCfg(**c["cfg"])
    m = build_module(c["ctor"], eo.synth_state_dict(cfg, c["wseed"]), loss_type="CFM-OT", device=DEV)
    x = torch.from_numpy(arr["ot_small_x"]).to(DEV)
    mask = torch.from_numpy(arr["ot_small_mask"]).to(DEV)
    torch.manual_seed(1); np.random.seed(1)
    out = m.training_step((x.clone(), mask, None), 0)
    assert torch.isfinite(out["loss"]) and out["loss"].requires_grad


# ------------------------------------------------------------------------------------------------ f3
def _jet_module(meta):
    from particle_fm_b200.models.flow_matching_no_sets import FLowMatchingNoSetsLitModule
    m = FLowMatchingNoSetsLitModule(optimizer=None, features=meta["features"], sigma=meta["sigma"], activation=meta["activation"],
                                    freqs=meta["freqs"])
    sd = no.synth_mlp_state_dict(meta["features"], meta["freqs"], meta["wseed"])
    m.flows[0].net.load_state_dict(sd, strict=True)
    return m.to(DEV), sd


def test_jet_feature_flow_vs_reference_golden():
    arr, meta = _load("next_jetflow")
    m, sd = _jet_module(meta)
    x, cond = torch.from_numpy(arr["x"]).to(DEV), torch.from_numpy(arr["cond"]).to(DEV)
    cnf = m.flows[0]
    with torch.no_grad():
        v_s = cnf(torch.from_numpy(arr["t_sample"]).to(DEV), x, cond=cond).cpu()
        v_t = cnf(torch.from_numpy(arr["t_train"]).to(DEV), x, cond=cond).cpu()
    assert rel_l2(v_s, torch.from_numpy(arr["v_sample"])) < 1e-5
    assert rel_l2(v_t, torch.from_numpy(arr["v_train"])) < 1e-5
    s = cnf.decode(torch.from_numpy(arr["z"]).to(DEV), cond, ode_solver="midpoint", ode_steps=meta["ode_steps"]).cpu()
    assert rel_l2(s, torch.from_numpy(arr["sample"])) < 1e-4
    torch.manual_seed(99)                                          # sample(): z from the CPU generator like the reference
    s2 = m.sample(x.shape[0], cond=cond, ode_solver="midpoint", ode_steps=meta["ode_steps"]).cpu()
    assert rel_l2(s2, torch.from_numpy(arr["sample"])) < 1e-4
    # training (torch ops on the device over the same parameters)
    loss = m.loss(x, cond=cond, draws=(torch.from_numpy(arr["loss_t"]).to(DEV), torch.from_numpy(arr["loss_z"]).to(DEV)))
    loss.backward()
    assert abs(float(loss) - float(arr["loss"])) <= 1e-5 * float(arr["loss"])
    for (k, p), want in zip(cnf.net.named_parameters(), arr["grad_norms"]):
        assert abs(float(p.grad.norm()) - want) <= 1e-3 * max(want, 1e-6), k


def test_jet_feature_flow_long_integration_and_ragged_tiles():
    """midpoint ode_steps=100 (the reference's setting) on a batch that does not fill the last 32-row tile, and Euler."""
    arr, meta = _load("next_jetflow")
    m, sd = _jet_module(meta)
    rs = np.random.RandomState(3)
    for B, solver, steps in ((70, "midpoint", 100), (5, "euler", 20)):
        z = torch.from_numpy(rs.normal(size=(B, meta["features"])).astype("float32"))
        cond = torch.from_numpy(rs.normal(size=(B, 1)).astype("float32"))
        with torch.no_grad():
            want = no.mlp_flow_sample(sd, z, cond, meta["freqs"], meta["activation"], steps, solver)
        got = m.flows[0].decode(z.to(DEV), cond.to(DEV), ode_solver=solver, ode_steps=steps).cpu()
        assert rel_l2(got, want) < 1e-4, (B, solver)


def test_lhco_chain_runs_on_device():
    """Step 1 (jet features | m_jj) feeds step 2 (particles | jet features) without leaving the GPU."""
    from particle_fm_b200.launch.lhco_chain import generate_lhco_chain
    arr, meta = _load("next_jetflow")
    jet, _ = _jet_module(meta)
    ctor = dict(features=3, hidden_dim=40, num_particles=37, frequencies=16, layers=2, latent=24, t_emb="cosine", t_local_cat=True,
                t_global_cat=True, add_time_to_input=False, global_cond_dim=4, local_cond_dim=4)
    cfg = eo.EpicCfg(feats=3, input_dim=3, hid=40, latent=24, layers=2, t_dim=32, t_local_cat=True, t_global_cat=True,
                     global_cond_dim=4, local_cond_dim=4)
    part = build_module(ctor, eo.synth_state_dict(cfg, 8), device=DEV)
    n = 50
    mjj = torch.linspace(2500, 4500, n)
    jm = [1200, 0, 0, 150, 20, 900, 0, 0, 100, 15]
    js = [300, 1, 1.8, 80, 8, 250, 1, 1.8, 60, 6]
    res = generate_lhco_chain(jet, part, mjj, jet_means=jm, jet_stds=js, mjj_mean=3500.0, mjj_std=600.0,
                              cond_means=[1000, 0, 0, 120], cond_stds=[300, 1, 1.8, 70], batch_size=16, jet_ode_steps=10,
                              ode_steps=6, device=DEV)
    data, mask = res["particle_data"], res["mask"].cpu().numpy()
    assert data.shape == (n, 2, 37, 3) and np.isfinite(data).all()
    assert res["jet_features"].is_cuda and tuple(res["jet_features"].shape) == (n, 2, 5)
    assert (data[mask[..., 0] == 0] == 0).all()                       # padded slots are zero
    mult = res["jet_features"][..., 4].cpu().numpy()
    assert np.array_equal(mask[..., 0].sum(-1), mult)


# ------------------------------------------------------------------------------------------------ f4
def _diff_module(meta, crit="huber"):
    cfg = eo.EpicCfg(**meta["cfg"])
    sd = eo.synth_state_dict(cfg, meta["wseed"])
    from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
    from helpers import full_state_dict
    m = SetFlowMatchingLitModule(optimizer=None, loss_type="diffusion", sigma=meta["sigma"], criterion=crit, **meta["ctor"])
    m.load_state_dict(full_state_dict(m, sd), strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("crit", ["huber", "mse"])
def test_diffusion_loss_vs_reference_golden(crit):
    arr, meta = _load("next_diffusion")
    m = _diff_module(meta, crit)
    x, mask = torch.from_numpy(arr["x"]).to(DEV), torch.from_numpy(arr["mask"]).to(DEV)
    draws = (torch.from_numpy(arr[f"loss_{crit}_t"]).to(DEV), torch.from_numpy(arr[f"loss_{crit}_z"]).to(DEV))
    loss = m.loss(x, mask=mask, cond=None, draws=draws)
    loss.backward()
    ref = float(arr[f"loss_{crit}"])
    assert abs(float(loss) - ref) <= 1e-5 * abs(ref), (float(loss), ref)
    grads = [p.grad for k, p in m.named_parameters() if k.startswith("flows.0.net.")]
    for gr, want in zip(grads, arr[f"loss_{crit}_grad_norms"]):
        assert abs(float(gr.norm()) - want) <= 1e-4 * max(want, 1e-6)


def test_diffusion_samplers_vs_reference_golden():
    arr, meta = _load("next_diffusion")
    m = _diff_module(meta)
    mask = torch.from_numpy(arr["mask"]).to(DEV)
    cnf = m.flows[0]
    s = cnf.decode(torch.from_numpy(arr["z_ddim"]).to(DEV), None, mask, ode_solver="ddim", ode_steps=meta["ddim_steps"]).cpu()
    assert rel_l2(s, torch.from_numpy(arr["sample_ddim"])) < 1e-4
    for solver, steps in meta["pf"]:
        s = cnf.decode(torch.from_numpy(arr[f"z_pf_{solver}{steps}"]).to(DEV), None, mask, ode_solver=solver, ode_steps=steps).cpu()
        assert rel_l2(s, torch.from_numpy(arr[f"sample_pf_{solver}{steps}"])) < 1e-4, solver
    # Euler-Maruyama with the reference's recorded per-step noise (the module draws its own with torch.randn_like)
    from particle_fm_b200.models.flow_matching_module import _diffusion_program
    t_eval, coef, _ = _diffusion_program(cnf, "em", meta["em_steps"])
    eng = cnf.net.engine(force_sync=True)
    codes = cnf.time_code(t_eval)
    got = eng.sample_diffusion(torch.from_numpy(arr["z_em"]).to(DEV), mask, None, codes, None, coef, "em",
                               noise=torch.from_numpy(arr["noise_em"]).to(DEV)).cpu()
    want = torch.from_numpy(arr["sample_em"]) * torch.from_numpy(arr["mask"])     # the reference leaves noise in padded slots
    assert rel_l2(got, want) < 1e-4
    torch.manual_seed(3)
    s = m.sample(6, mask=mask, ode_solver="em", ode_steps=5)
    assert torch.isfinite(s).all() and float(s[mask.expand_as(s) == 0].abs().max()) == 0.0
