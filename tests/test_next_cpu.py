"""CPU: the oracle restatements of the "next" rows against the golden vectors generated from the
unmodified reference (oracle/make_golden_next.py), and the host-side logic of those rows."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import next_oracle as no

from helpers import GOLDEN_DIR


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files if k != "meta"}, json.loads(str(z["meta"]))


def _bounds(n, batch):
    b = [(i * batch, (i + 1) * batch) for i in range(n // batch)]
    if n % batch:
        b.append((n - n % batch, n))
    return b


def test_post_processing_oracle_vs_reference_golden():
    arr, meta = _load("next_post")
    for name, c in meta.items():
        kw = c["kw"]
        raw, mask = torch.from_numpy(arr[f"{name}_raw"]), torch.from_numpy(arr[f"{name}_mask"])
        parts = [no.post_process(raw[lo:hi], mask[lo:hi], kw.get("normalized_data", False), kw.get("normalize_sigma", 5),
                                 arr[f"{name}_means"], arr[f"{name}_stds"], kw.get("log_pt", False),
                                 kw.get("pt_standardization", False), kw.get("variable_set_sizes", False))
                 for lo, hi in _bounds(c["n"], c["batch"])]
        assert np.array_equal(torch.cat(parts).numpy(), arr[f"{name}_out"]), name


def test_post_coefficients_match_the_eager_arithmetic():
    from particle_fm_b200.utils.data_generation import post_coefficients
    arr, meta = _load("next_post")
    for name, c in meta.items():
        kw = c["kw"]
        sc, sh, lc, fo = post_coefficients(c["F"], kw.get("normalized_data", False), kw.get("normalize_sigma", 5), arr[f"{name}_means"],
                                       arr[f"{name}_stds"], kw.get("log_pt", False), kw.get("pt_standardization", False))
        if not kw.get("normalized_data", False):
            assert sc is None and lc == -1
            continue
        raw = torch.from_numpy(arr[f"{name}_raw"])
        v = raw * torch.tensor(sc, dtype=torch.float32) + torch.tensor(sh, dtype=torch.float32)      # mul, then add: two roundings
        if fo >= 0:
            v[:, 1:, fo] = raw[:, 1:, fo]                                                            # the reference's particle-0 quirk
        if lc >= 0:
            v[..., lc] = 1.0 - torch.exp(v[..., lc])
        if kw.get("variable_set_sizes", False):
            v = v * torch.from_numpy(arr[f"{name}_mask"])
        np.testing.assert_allclose(v.numpy(), arr[f"{name}_out"], rtol=2e-6, atol=1e-7)
        cols = [f for f in range(c["F"]) if f != lc]
        assert np.array_equal(v.numpy()[..., cols], arr[f"{name}_out"][..., cols]), name       # affine part: bit-equal


def test_jet_feature_flow_oracle_vs_reference_golden():
    arr, meta = _load("next_jetflow")
    F, fr, act = meta["features"], meta["freqs"], meta["activation"]
    sd = no.synth_mlp_state_dict(F, fr, meta["wseed"])
    x, cond = torch.from_numpy(arr["x"]), torch.from_numpy(arr["cond"])
    with torch.no_grad():
        assert np.array_equal(no.mlp_flow_forward(sd, torch.from_numpy(arr["t_sample"]), x, cond, fr, act).numpy(), arr["v_sample"])
        assert np.array_equal(no.mlp_flow_forward(sd, torch.from_numpy(arr["t_train"]), x, cond, fr, act).numpy(), arr["v_train"])
        s = no.mlp_flow_sample(sd, torch.from_numpy(arr["z"]), cond, fr, act, meta["ode_steps"])
    assert np.array_equal(s.numpy(), arr["sample"])
    loss = no.mlp_flow_loss(sd, x, cond, torch.from_numpy(arr["loss_t"]), torch.from_numpy(arr["loss_z"]), fr, act, meta["sigma"])
    assert float(loss) == float(arr["loss"])


def test_jet_feature_module_mirrors_the_reference_state_dict():
    from particle_fm_b200.models.flow_matching_no_sets import FLowMatchingNoSetsLitModule
    arr, meta = _load("next_jetflow")
    m = FLowMatchingNoSetsLitModule(optimizer=None, features=meta["features"], freqs=meta["freqs"], activation=meta["activation"])
    names = [k for k, _ in m.flows[0].net.named_parameters()]
    assert names == [str(s) for s in arr["grad_names"]]
    sd = no.synth_mlp_state_dict(meta["features"], meta["freqs"], meta["wseed"])
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(p.shape) for k, p in m.flows[0].net.named_parameters()}
    assert torch.allclose(m.flows[0].freqs, torch.arange(1, meta["freqs"] + 1) * torch.pi)
    with pytest.raises(RuntimeError, match="CUDA"):                      # no CPU fallback
        m.sample(4, cond=torch.zeros(4, 1))


def test_cfm_ot_oracle_vs_reference_golden():
    arr, meta = _load("next_cfmot")
    for name, c in meta.items():
        cfg = eo.EpicCfg(**c["cfg"])
        sd = eo.synth_state_dict(cfg, c["wseed"])
        ctor = c["ctor"]
        kw = dict(t_emb=ctor["t_emb"], frequencies=ctor["frequencies"], add_time_to_input=ctor["add_time_to_input"])
        g = lambda k: torch.from_numpy(arr[f"{name}_{k}"])
        # coupling + interpolation, pinned against the reference's own forward (network replaced by a recorder)
        loss_s, x1, _, y = no.cfm_ot_loss(lambda tt, yy, mk: torch.zeros_like(yy), g("x"), g("mask"), g("x0"), g("t"),
                                          arr[f"{name}_u"], g("eps"), c["sigma"], mask_mode="reference")
        assert np.array_equal(x1.numpy(), arr[f"{name}_x1_after"]) and np.array_equal(y.numpy(), arr[f"{name}_y"])
        assert float(loss_s) == float(arr[f"{name}_loss_stub"])
        with torch.no_grad():
            vf = lambda tt, yy, mk: eo.cnf_forward(sd, cfg, tt, yy, None, mk, **kw)
            loss, _, mask_ot, _ = no.cfm_ot_loss(vf, g("x"), g("mask"), g("x0"), g("t"), arr[f"{name}_u"], g("eps"), c["sigma"])
        assert float(loss) == float(arr[f"{name}_loss"])
        assert np.array_equal(mask_ot.numpy(), arr[f"{name}_mask_ot"])


def test_pair_draws_follow_numpy_choice():
    """picks_from_uniform restates RandomState.choice(N*N, p=plan, size=N): same seed -> same row picks."""
    from particle_fm_b200.models.components.losses import ConditionalFlowMatchingOTLoss as L
    N = 23
    perm = np.random.RandomState(3).permutation(N)
    pi = np.zeros((N, N)); pi[np.arange(N), perm] = 1.0 / N
    p = pi.flatten() / pi.sum()
    np.random.seed(99)
    choices = np.random.choice(N * N, p=p, size=N)
    i_ref, j_ref = np.divmod(choices, N)
    np.random.seed(99)
    u = np.random.random_sample((1, N))
    picks = L.picks_from_uniform(u)[0].numpy()
    assert np.array_equal(picks, i_ref) and np.array_equal(perm[picks], j_ref)
    assert np.array_equal(no.choice_from_uniform(p, u[0]), choices)


def test_diffusion_oracle_vs_reference_golden():
    arr, meta = _load("next_diffusion")
    ctor = dict(meta["ctor"]); dc = ctor.pop("diff_config")
    cfg = eo.EpicCfg(**meta["cfg"])
    sd = eo.synth_state_dict(cfg, meta["wseed"])
    kw = dict(t_emb=ctor["t_emb"], frequencies=ctor["frequencies"], add_time_to_input=ctor["add_time_to_input"])
    x, mask = torch.from_numpy(arr["x"]), torch.from_numpy(arr["mask"])
    vf = lambda tt, y: eo.cnf_forward(sd, cfg, tt, y, None, mask, **kw)
    with torch.no_grad():
        for crit in ("huber", "mse"):
            loss = no.diffusion_loss(vf, x, mask, torch.from_numpy(arr[f"loss_{crit}_t"]), torch.from_numpy(arr[f"loss_{crit}_z"]), dc, crit)
            assert float(loss) == float(arr[f"loss_{crit}"]), crit
        assert np.array_equal(no.ddim_sample(vf, torch.from_numpy(arr["z_ddim"]), meta["ddim_steps"], dc).numpy(), arr["sample_ddim"])
        noise = list(torch.from_numpy(arr["noise_em"]))
        assert np.array_equal(no.em_sample(vf, torch.from_numpy(arr["z_em"]), meta["em_steps"], dc, noise).numpy(), arr["sample_em"])
        for solver, steps in meta["pf"]:
            s = no.pf_ode_sample(vf, torch.from_numpy(arr[f"z_pf_{solver}{steps}"]), steps, solver, dc)
            assert np.array_equal(s.numpy(), arr[f"sample_pf_{solver}{steps}"])


def test_diffusion_step_programs_match_the_oracle_schedule():
    """The host-computed coefficient tables handed to pfm_epic_sample_diffusion equal the values the oracle's samplers
    use step by step (same fp32 recurrences)."""
    from particle_fm_b200.models.flow_matching_module import CNF, _diffusion_program, fixed_step_grid
    dc = {"max_sr": 0.999, "min_sr": 0.02}
    cnf = CNF(features=3, hidden_dim=16, num_particles=5, frequencies=4, layers=1, latent=4, t_emb="cosine", t_local_cat=True,
              t_global_cat=True, add_time_to_input=False, loss_type="diffusion", diff_config=dc)
    n = 7
    t, coef, dt = _diffusion_program(cnf, "ddim", n)
    tm = torch.ones(1)
    for s in range(n):
        sr, nr = no.diff_rates(tm.view(-1, 1, 1), **dc)
        assert float(t[s]) == float(tm[0]) and float(coef[s, 0]) == float(sr) and float(coef[s, 1]) == float(nr)
        tm = tm - 1 / n
        nsr, nnr = no.diff_rates(tm.view(-1, 1, 1), **dc)
        assert float(coef[s, 2]) == float(nsr) and float(coef[s, 3]) == float(nnr)
    t, coef, dt = _diffusion_program(cnf, "em", n)
    tm = torch.ones(1)
    for s in range(n):
        b = no.diff_betas(tm.view(-1, 1, 1), **dc)
        assert float(t[s]) == float(tm[0]) and float(coef[s, 0]) == float(b) and float(coef[s, 3]) == float((b * (1 / n)).sqrt())
        tm = tm - 1 / n
    t, coef, dt = _diffusion_program(cnf, "midpoint", n)
    te, dte = fixed_step_grid(n, "midpoint")
    assert torch.equal(t, te) and torch.equal(dt, dte) and coef.shape == (2 * (n - 1), 4)
    assert torch.equal(coef[:, 0], no.diff_betas(te.view(-1, 1, 1), **dc).reshape(-1))


def test_diffusion_and_cfmot_modules_construct():
    from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
    kw = dict(features=3, hidden_dim=16, num_particles=5, frequencies=4, layers=1, latent=4, t_emb="cosine", t_local_cat=True,
              t_global_cat=True, add_time_to_input=False)
    m = SetFlowMatchingLitModule(optimizer=None, loss_type="diffusion", diff_config={"max_sr": 0.999, "min_sr": 0.02}, **kw)
    assert type(m.loss).__name__ == "DiffusionLoss"
    m = SetFlowMatchingLitModule(optimizer=None, loss_type="CFM-OT", **kw)
    assert type(m.loss).__name__ == "ConditionalFlowMatchingOTLoss"
    with pytest.raises(SyntaxError):                                     # flow_matching_module.py:328-329
        m.flows[0].decode(torch.zeros(1, 5, 3), None, None, ode_solver="ddim")
