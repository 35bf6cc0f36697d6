"""Oracle (test infrastructure): golden vectors for the scope table's "next" rows from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_next
Writes tests/golden/next_*.npz and asserts that oracle/next_oracle.py reproduces every reference output:
  next_post.npz       generate_data (utils/data_generation.py:17-174) driven with a stub model whose sample() returns
                      recorded tensors -> the reference's own batching + post-processing
  next_jetflow.npz    FLowMatchingNoSetsLitModule (models/flow_matching_no_sets.py): vector field, sample(), loss + grads
  next_cfmot.npz      ConditionalFlowMatchingOTLoss.forward (losses.py:147-204) with scipy's assignment solver standing
                      in for the absent POT ``ot.emd`` (parity unpinned for the plan itself, see oracle/next_oracle.py)
  next_diffusion.npz  DiffusionLoss.forward, ddim_sampler, euler_maruyama_sampler, probability-flow decode
"""
from __future__ import annotations

import importlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

from . import epic_oracle as eo
from . import next_oracle as no
from . import ref_shim
from .make_golden import GOLDEN_DIR, SIGMA, build_reference, cfg_from_ctor


def _load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference_generate_data():
    """particle_fm/utils/data_generation.py with its one project import (data/components/utils.py) loaded by path;
    energyflow (absent, unused on this path) is an empty stand-in."""
    ref_shim.load()
    root = ref_shim.REF_ROOT
    if "energyflow" not in sys.modules:
        sys.modules["energyflow"] = types.ModuleType("energyflow")
    for pkg in ("particle_fm.data", "particle_fm.data.components"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    _load_file("particle_fm.data.components.utils", os.path.join(root, "particle_fm", "data", "components", "utils.py"))
    return _load_file("particle_fm.utils.data_generation", os.path.join(root, "particle_fm", "utils", "data_generation.py"))


class _StubModel:
    """What generate_data needs from a model: .to() and .sample(); returns the recorded raw samples in call order."""

    def __init__(self, raw):
        self.raw, self.pos = raw, 0

    def to(self, device):
        return self

    def sample(self, n_samples, cond=None, mask=None, ode_solver="midpoint", ode_steps=100):
        out = self.raw[self.pos:self.pos + n_samples].clone()
        self.pos += n_samples
        return out


POST_CASES = {
    "plain": dict(n=23, N=30, F=3, batch=8, kw=dict(normalized_data=True, normalize_sigma=5, variable_set_sizes=True)),
    "ptstd_logpt": dict(n=16, N=20, F=3, batch=16, kw=dict(normalized_data=True, normalize_sigma=5, log_pt=True,
                                                           pt_standardization=True, variable_set_sizes=True)),
    "logpt_nomask": dict(n=10, N=12, F=3, batch=4, kw=dict(normalized_data=True, normalize_sigma=3, log_pt=True)),
    "raw_masked": dict(n=9, N=7, F=4, batch=4, kw=dict(normalized_data=False, variable_set_sizes=True)),
    "feat8": dict(n=12, N=16, F=8, batch=5, kw=dict(normalized_data=True, normalize_sigma=5, variable_set_sizes=True)),
}


def golden_post():
    gd = load_reference_generate_data()
    out = {}
    for name, c in POST_CASES.items():
        rs = np.random.RandomState(hash(name) % 1000 + 11)
        n, N, F = c["n"], c["N"], c["F"]
        raw = torch.from_numpy(rs.normal(0, 2.0, size=(n, N, F)).astype("float32"))
        n_real = rs.randint(1, N + 1, size=n)
        mask = torch.from_numpy((np.arange(N)[None, :] < n_real[:, None]).astype("float32")).unsqueeze(-1)
        means = rs.normal(0, 1, size=F).astype("float32")
        stds = rs.uniform(0.5, 2.0, size=F).astype("float32")
        kw = dict(c["kw"])
        use_mask = kw.get("variable_set_sizes", False)
        data, _ = gd.generate_data(_StubModel(raw), n, batch_size=c["batch"], device="cpu", mask=mask if use_mask else None,
                                   means=means, stds=stds, verbose=False, **kw)
        # the oracle's restatement, batch by batch like the reference
        parts = []
        bounds = [(i * c["batch"], (i + 1) * c["batch"]) for i in range(n // c["batch"])]
        if n % c["batch"]:
            bounds.append((n - n % c["batch"], n))
        for lo, hi in bounds:
            parts.append(no.post_process(raw[lo:hi], mask[lo:hi], kw.get("normalized_data", False), kw.get("normalize_sigma", 5),
                                         means, stds, kw.get("log_pt", False), kw.get("pt_standardization", False), use_mask))
        mine = torch.cat(parts).numpy()
        assert np.array_equal(mine, data), name
        out[f"{name}_raw"] = raw.numpy(); out[f"{name}_mask"] = mask.numpy()
        out[f"{name}_means"] = means; out[f"{name}_stds"] = stds; out[f"{name}_out"] = data
    out["meta"] = np.array(json.dumps({k: dict(n=c["n"], N=c["N"], F=c["F"], batch=c["batch"], kw=c["kw"]) for k, c in POST_CASES.items()}))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "next_post.npz"), **out)
    print("next_post: ok")


JET = dict(features=10, freqs=3, activation="ELU", wseed=31, B=37, sigma=1e-4, ode_steps=12)


def golden_jetflow():
    ref_shim.load()
    fns = importlib.import_module("particle_fm.models.flow_matching_no_sets")
    F, freqs = JET["features"], JET["freqs"]
    sd = no.synth_mlp_state_dict(F, freqs, JET["wseed"])
    m = fns.FLowMatchingNoSetsLitModule(optimizer=None, features=F, sigma=JET["sigma"], activation=JET["activation"], freqs=freqs)
    full = {"flows.0.freqs": m.flows[0].freqs.clone(), "loss.flows.0.freqs": m.flows[0].freqs.clone()}
    for pre in ("flows.0.net.", "loss.flows.0.net."):
        full.update({pre + k: v.clone() for k, v in sd.items()})
    m.load_state_dict(full, strict=True)
    rs = np.random.RandomState(5)
    B = JET["B"]
    x = torch.from_numpy(rs.normal(0, 1, size=(B, F)).astype("float32"))
    cond = torch.from_numpy(rs.normal(0, 1, size=(B, 1)).astype("float32"))
    t0 = torch.tensor(0.3127, dtype=torch.float32)
    tr = torch.from_numpy(rs.uniform(0, 1, size=B).astype("float32"))
    cnf = m.flows[0]
    out = dict(x=x.numpy(), cond=cond.numpy(), t_sample=t0.numpy(), t_train=tr.numpy())
    with torch.no_grad():
        v_s = cnf(t0, x, cond=cond)
        v_t = cnf(tr, x, cond=cond)
        assert torch.equal(v_s, no.mlp_flow_forward(sd, t0, x, cond, freqs, JET["activation"]))
        assert torch.equal(v_t, no.mlp_flow_forward(sd, tr, x, cond, freqs, JET["activation"]))
    out.update(v_sample=v_s.numpy(), v_train=v_t.numpy())
    torch.manual_seed(99)
    s_r = m.sample(B, cond=cond, ode_solver="midpoint", ode_steps=JET["ode_steps"])
    torch.manual_seed(99)
    z = torch.randn(B, F)
    with torch.no_grad():
        s_o = no.mlp_flow_sample(sd, z, cond, freqs, JET["activation"], JET["ode_steps"])
    assert torch.equal(s_r, s_o), (s_r - s_o).abs().max()
    out.update(z=z.numpy(), sample=s_r.numpy())
    torch.manual_seed(7)
    loss_r = m.loss(x, cond=cond)
    loss_r.backward()
    torch.manual_seed(7)
    t = torch.rand_like(x[..., 0]).unsqueeze(-1)
    zz = torch.randn_like(x)
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_o = no.mlp_flow_loss(sd_g, x, cond, t, zz, freqs, JET["activation"], JET["sigma"])
    loss_o.backward()
    assert torch.equal(loss_r.detach(), loss_o.detach())
    names = [k for k, _ in cnf.net.named_parameters()]
    for k, p in cnf.net.named_parameters():
        assert torch.allclose(p.grad, sd_g[k].grad, rtol=1e-5, atol=1e-7), k
    out.update(loss=loss_r.detach().numpy(), loss_t=t.numpy(), loss_z=zz.numpy(), grad_names=np.array(names),
               grad_norms=np.array([float(p.grad.norm()) for _, p in cnf.net.named_parameters()], dtype="float64"))
    out["meta"] = np.array(json.dumps(JET))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "next_jetflow.npz"), **out)
    print(f"next_jetflow: ok  loss={float(loss_r):.6f}")


OT_CASES = {
    # JetNet-30-like: ragged prefix masks, zero padding participates in the transport like in the reference
    "ot_c1": dict(ctor=dict(features=3, hidden_dim=128, num_particles=30, frequencies=16, layers=6, latent=10, t_emb="cosine",
                            t_local_cat=True, t_global_cat=True, add_time_to_input=False), B=6, wseed=12345, seed=201),
    "ot_small": dict(ctor=dict(features=3, hidden_dim=48, num_particles=17, frequencies=6, layers=2, latent=16, t_emb="sincos",
                               t_local_cat=False, t_global_cat=False, add_time_to_input=True), B=5, wseed=7, seed=202),
}


def golden_cfmot():
    ref = ref_shim.load()
    pot = sys.modules["ot"]
    pot.unif = lambda n: np.ones(n) / n
    pot.emd = lambda a, b, M: no.emd_uniform(np.asarray(M))
    out, meta = {}, {}
    for name, case in OT_CASES.items():
        ctor = case["ctor"]
        cfg = cfg_from_ctor(ctor)
        sd = eo.synth_state_dict(cfg, case["wseed"])
        N, Fd, B = ctor["num_particles"], ctor["features"], case["B"]
        x, mask, _ = eo.synth_cloud(B, N, Fd, case["seed"])
        kw = dict(t_emb=ctor["t_emb"], frequencies=ctor["frequencies"], add_time_to_input=ctor["add_time_to_input"])
        # The reference's forward cannot run with its own networks: mask_ot is left as the LAST jet's (N, 1) mask (:189)
        # and EPiC_encoder then fails at epic.py:370 ((B, H) / (N,)).  Pin the coupling + interpolation against the
        # reference's code with a recording stand-in for the flow, then record the per-jet-mask loss of the oracle.
        class _Rec(torch.nn.Module):
            def forward(self, t, y, mask=None, cond=None):
                self.t, self.y, self.mask = t.clone(), y.clone(), mask.clone()
                return torch.zeros_like(y)
        rec = _Rec()
        loss_mod = ref.losses.ConditionalFlowMatchingOTLoss(flows=torch.nn.ModuleList([rec]), sigma=SIGMA)
        torch.manual_seed(31337)
        np.random.seed(4711)
        x_ref = x.clone()
        loss_r = loss_mod(x_ref, mask=mask, cond=None)          # re-indexes x_ref in place (x1 = x aliases it)
        # same draws, in the reference's order: randn_like(x) | rand(B) | N uniforms per jet (np.random.choice) | randn_like
        torch.manual_seed(31337)
        np.random.seed(4711)
        x0 = torch.randn_like(x)
        t = torch.rand_like(torch.ones(B))
        u = np.random.random_sample((B, N))
        eps = torch.randn_like(x0)
        zero_vf = lambda tt, y, mk: torch.zeros_like(y)
        loss_s, x1_o, mask_last, y_o = no.cfm_ot_loss(zero_vf, x, mask, x0, t, u, eps, SIGMA, mask_mode="reference")
        assert torch.equal(loss_r, loss_s), (name, loss_r, loss_s)
        assert torch.equal(x_ref, x1_o) and torch.equal(rec.y, y_o) and torch.equal(rec.mask, mask_last), name
        sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        vf = lambda tt, y, mk: eo.cnf_forward(sd_g, cfg, tt, y, None, mk, **kw)
        loss_o, _, mask_ot, _ = no.cfm_ot_loss(vf, x, mask, x0, t, u, eps, SIGMA, mask_mode="per_jet")
        loss_o.backward()
        names = [k for k in sd_g]
        out.update({f"{name}_x": x.numpy(), f"{name}_mask": mask.numpy(), f"{name}_x0": x0.numpy(), f"{name}_t": t.numpy(),
                    f"{name}_u": u, f"{name}_eps": eps.numpy(), f"{name}_loss_stub": loss_r.detach().numpy(),
                    f"{name}_y": rec.y.numpy(), f"{name}_x1_after": x_ref.numpy(), f"{name}_mask_ot": mask_ot.numpy(),
                    f"{name}_loss": loss_o.detach().numpy(), f"{name}_grad_names": np.array(names),
                    f"{name}_grad_norms": np.array([float(sd_g[k].grad.norm()) for k in names], dtype="float64")})
        loss_r = loss_o
        meta[name] = dict(ctor=ctor, cfg=cfg.as_dict(), wseed=case["wseed"], B=B, sigma=SIGMA)
        print(f"  {name}: loss = {float(loss_r):.6f}")
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "next_cfmot.npz"), **out)
    print("next_cfmot: ok")


DIFF = dict(ctor=dict(features=3, hidden_dim=128, num_particles=30, frequencies=16, layers=6, latent=10, t_emb="cosine",
                      t_local_cat=True, t_global_cat=True, add_time_to_input=False,
                      diff_config={"max_sr": 0.999, "min_sr": 0.02}),
            B=6, wseed=12345, seed=301, ddim_steps=9, em_steps=7, pf=[("euler", 8), ("midpoint", 6)])


def golden_diffusion():
    ref = ref_shim.load()
    ctor = DIFF["ctor"]
    dc = ctor["diff_config"]
    cfg = cfg_from_ctor(ctor)
    sd = eo.synth_state_dict(cfg, DIFF["wseed"])
    N, Fd, B = ctor["num_particles"], ctor["features"], DIFF["B"]
    x, mask, _ = eo.synth_cloud(B, N, Fd, DIFF["seed"])
    kw = dict(t_emb=ctor["t_emb"], frequencies=ctor["frequencies"], add_time_to_input=ctor["add_time_to_input"])
    out = dict(x=x.numpy(), mask=mask.numpy())
    # ---- loss (both criteria)
    for crit in ("huber", "mse"):
        m = ref.fm.SetFlowMatchingLitModule(optimizer=None, loss_type="diffusion", sigma=SIGMA, criterion=crit, **ctor)
        full = {}
        for pre in ("flows.0.", "loss.flows.0."):
            full.update({(pre + "net." + k): v.clone() for k, v in sd.items()})
            full[pre + "frequencies"] = m.flows[0].frequencies.clone()
        m.load_state_dict(full, strict=True)
        torch.manual_seed(2024)
        loss_r = m.loss(x, mask=mask, cond=None)
        loss_r.backward()
        torch.manual_seed(2024)
        t = torch.rand_like(torch.ones(B))
        z = torch.randn_like(x)
        sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        vf = lambda tt, y: eo.cnf_forward(sd_g, cfg, tt, y, None, mask, **kw)
        loss_o = no.diffusion_loss(vf, x, mask, t, z, dc, crit)
        loss_o.backward()
        assert torch.equal(loss_r.detach(), loss_o.detach()), (crit, loss_r, loss_o)
        params = list(m.flows[0].net.named_parameters())
        for k, p in params:
            assert torch.allclose(p.grad, sd_g[k].grad, rtol=1e-5, atol=1e-7), (crit, k)
        out.update({f"loss_{crit}": loss_r.detach().numpy(), f"loss_{crit}_t": t.numpy(), f"loss_{crit}_z": z.numpy(),
                    f"loss_{crit}_grad_norms": np.array([float(p.grad.norm()) for _, p in params], dtype="float64")})
        print(f"  diffusion loss[{crit}] = {float(loss_r):.6f}")
    # ---- samplers through the reference's sample() / decode()
    vf0 = lambda tt, y: eo.cnf_forward(sd, cfg, tt, y, None, mask, **kw)
    torch.manual_seed(55)
    s_r = m.sample(B, mask=mask, ode_solver="ddim", ode_steps=DIFF["ddim_steps"])
    torch.manual_seed(55)
    z = torch.randn(B, N, Fd) * mask
    with torch.no_grad():
        s_o = no.ddim_sample(vf0, z, DIFF["ddim_steps"], dc)
    assert torch.equal(s_r, s_o), (s_r - s_o).abs().max()
    out.update(z_ddim=(z / 1).numpy(), sample_ddim=s_r.numpy())
    torch.manual_seed(56)
    s_r = m.sample(B, mask=mask, ode_solver="em", ode_steps=DIFF["em_steps"])
    torch.manual_seed(56)
    z = torch.randn(B, N, Fd) * mask
    noise = []
    with torch.no_grad():
        # the reference draws randn_like(x_t) inside the loop; the net itself consumes no random numbers
        noise = [torch.randn_like(z) for _ in range(DIFF["em_steps"])]
        s_o = no.em_sample(vf0, z, DIFF["em_steps"], dc, noise)
    assert torch.equal(s_r, s_o), (s_r - s_o).abs().max()
    out.update(z_em=z.numpy(), noise_em=torch.stack(noise).numpy(), sample_em=s_r.numpy())
    for solver, steps in DIFF["pf"]:
        torch.manual_seed(57)
        s_r = m.sample(B, mask=mask, ode_solver=solver, ode_steps=steps)
        torch.manual_seed(57)
        z = torch.randn(B, N, Fd) * mask
        with torch.no_grad():
            s_o = no.pf_ode_sample(vf0, z, steps, solver, dc)
        assert torch.equal(s_r, s_o), (solver, (s_r - s_o).abs().max())
        out.update({f"z_pf_{solver}{steps}": z.numpy(), f"sample_pf_{solver}{steps}": s_r.numpy()})
    meta = dict(ctor=ctor, cfg=cfg.as_dict(), wseed=DIFF["wseed"], B=B, sigma=SIGMA, ddim_steps=DIFF["ddim_steps"],
                em_steps=DIFF["em_steps"], pf=DIFF["pf"])
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "next_diffusion.npz"), **out)
    print("next_diffusion: ok")


def main():
    torch.set_num_threads(1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    golden_post()
    golden_jetflow()
    golden_cfmot()
    golden_diffusion()


if __name__ == "__main__":
    sys.exit(main())
