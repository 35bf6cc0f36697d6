"""Oracle (test infrastructure): generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
For every case it
  1. builds the reference's own SetFlowMatchingLitModule (through oracle/ref_shim.py),
  2. loads deterministic synthetic weights (numpy RandomState stream -> reproducible in tests
     without the reference),
  3. records the reference's outputs: vector field in sampling mode (0-dim t) and training mode
     (per-jet t), the three flow-matching losses with their parameter gradients, and sample()
     end points (Euler / midpoint through the torchdyn stand-in),
  4. asserts that oracle/*.py reproduces each of them (bit-for-bit for the network and losses).
The .npz files hold inputs and reference outputs only (weights are regenerated from the seed).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import epic_oracle as eo
from . import loss_oracle as lo
from . import ode_oracle as oo
from . import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (module ctor kwargs, cloud spec)
CASES = {
    # configs/model/flow_matching.yaml:15-29 + experiment/jetnet/fm_tops30.yaml (BASELINE configs[0])
    "c1_jetnet30": dict(
        ctor=dict(features=3, hidden_dim=128, num_particles=30, frequencies=16, layers=6, latent=10,
                  t_emb="cosine", t_local_cat=True, t_global_cat=True, add_time_to_input=False),
        cloud=dict(B=6, seed=101), wseed=12345, ode=[("euler", 100), ("midpoint", 12)]),
    # BASELINE configs[1]: JetNet-150, same net (experiment/jetnet/fm_tops150.yaml)
    "c2_jetnet150": dict(
        ctor=dict(features=3, hidden_dim=128, num_particles=150, frequencies=16, layers=6, latent=10,
                  t_emb="cosine", t_local_cat=True, t_global_cat=True, add_time_to_input=False),
        cloud=dict(B=4, seed=102), wseed=12345, ode=[("midpoint", 9)]),
    # bare-constructor mode (flow_matching_module.py:381-412): sincos code on the input only
    "bare_sincos": dict(
        ctor=dict(features=3, hidden_dim=48, num_particles=17, frequencies=6, layers=2, latent=16,
                  t_emb="sincos", t_local_cat=False, t_global_cat=False, add_time_to_input=True),
        cloud=dict(B=5, seed=103), wseed=7, ode=[("euler", 7)]),
    # LHCO-like: global AND local conditioning, latent > hidden-ish, ragged masks
    # (experiment/lhco/both_jets.yaml:26-32 scaled down)
    "cond_lhco_like": dict(
        ctor=dict(features=3, hidden_dim=40, num_particles=37, frequencies=16, layers=2, latent=24,
                  t_emb="cosine", t_local_cat=True, t_global_cat=True, add_time_to_input=False,
                  global_cond_dim=4, local_cond_dim=4),
        cloud=dict(B=5, seed=104, cond_dim=4, ragged=True), wseed=8, ode=[("midpoint", 6)]),
    # JetClass-cond-like: 13 features, global cond only, odd hidden size
    # (experiment/jetclass_cond.yaml:32-39 scaled down)
    "cond_jetclass_like": dict(
        ctor=dict(features=13, hidden_dim=44, num_particles=21, frequencies=16, layers=3, latent=16,
                  t_emb="cosine", t_local_cat=True, t_global_cat=True, add_time_to_input=False,
                  global_cond_dim=12, local_cond_dim=0),
        cloud=dict(B=4, seed=105, cond_dim=12), wseed=9, ode=[("euler", 5)]),
    # time only on the global path, plain (non weight-normed) linears
    "tglobal_plain": dict(
        ctor=dict(features=4, hidden_dim=32, num_particles=9, frequencies=4, layers=1, latent=6,
                  t_emb="cosine", t_local_cat=False, t_global_cat=True, add_time_to_input=True,
                  wrapper_func="not_a_wrapper"),
        cloud=dict(B=3, seed=106), wseed=10, ode=[("midpoint", 4)]),
}

SIGMA = 1e-4


def cfg_from_ctor(c) -> eo.EpicCfg:
    T = 2 * c["frequencies"]
    return eo.EpicCfg(
        feats=c["features"], input_dim=c["features"] + (T if c["add_time_to_input"] else 0),
        hid=c["hidden_dim"], latent=c["latent"], layers=c["layers"], t_dim=T,
        t_local_cat=c["t_local_cat"], t_global_cat=c["t_global_cat"],
        global_cond_dim=c.get("global_cond_dim", 0), local_cond_dim=c.get("local_cond_dim", 0))


def build_reference(ref, ctor, sd, loss_type):
    m = ref.fm.SetFlowMatchingLitModule(optimizer=None, loss_type=loss_type, sigma=SIGMA, **ctor)
    # the loss module holds the same ModuleList, so every key appears twice in the reference's
    # state_dict: "flows.0.*" and "loss.flows.0.*" (flow_matching_module.py:447-466)
    full = {}
    for pre in ("flows.0.", "loss.flows.0."):
        full.update({(pre + "net." + k): v.clone() for k, v in sd.items()})
        full[pre + "frequencies"] = m.flows[0].frequencies.clone()
    m.load_state_dict(full, strict=True)
    return m


def main():
    ref = ref_shim.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)          # deterministic reduction order while pinning
    for name, case in CASES.items():
        ctor = case["ctor"]
        cfg = cfg_from_ctor(ctor)
        wn = ctor.get("wrapper_func", "weight_norm") == "weight_norm"
        sd = eo.synth_state_dict(cfg, case["wseed"], weight_norm=wn)
        cl = case["cloud"]
        N, Fd = ctor["num_particles"], ctor["features"]
        x, mask, cond = eo.synth_cloud(cl["B"], N, Fd, cl["seed"], cond_dim=cl.get("cond_dim", 0),
                                       ragged=cl.get("ragged", False))
        B = x.shape[0]
        out = {"x": x.numpy(), "mask": mask.numpy()}
        if cond is not None:
            out["cond"] = cond.numpy()
        kw = dict(t_emb=ctor["t_emb"], frequencies=ctor["frequencies"],
                  add_time_to_input=ctor["add_time_to_input"])
        vf_o = lambda t, y: eo.cnf_forward(sd, cfg, t, y, cond, mask, **kw)

        m = build_reference(ref, ctor, sd, "FM-OT")
        cnf = m.flows[0]
        # -- vector field, sampling mode (0-dim t) and training mode (per-jet t broadcast to (B,N))
        t0 = torch.tensor(0.7371, dtype=torch.float32)
        tj = torch.from_numpy(np.random.RandomState(cl["seed"] + 1).uniform(0, 1, B).astype("float32"))
        tbn = tj.unsqueeze(-1).repeat_interleave(N, dim=1)
        with torch.no_grad():
            v_s = cnf(t0, x, cond=cond, mask=mask)
            v_t = cnf(tbn, x, cond=cond, mask=mask)
            v_i = cnf(t0, x, cond=cond, mask=mask.long())          # int64 mask (jetnet_datamodule.py:247-249)
            assert torch.equal(v_s, vf_o(t0, x)), name
            assert torch.equal(v_t, vf_o(tbn, x)), name
            assert torch.equal(v_i, v_s), name
        out.update(t_sample=t0.numpy(), t_train=tj.numpy(), v_sample=v_s.numpy(), v_train=v_t.numpy())

        # -- losses with gradients (reference RNG order reproduced by lo.draw_loss_randoms)
        for kind in ("FM-OT", "CFM", "droid"):
            m = build_reference(ref, ctor, sd, kind)
            torch.manual_seed(4242)
            loss_r = m.loss(x, mask=mask, cond=cond)
            loss_r.backward()
            torch.manual_seed(4242)
            t, n0, n1 = lo.draw_loss_randoms(kind, x)
            sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            vf_g = lambda tt, y: eo.cnf_forward(sd_g, cfg, tt, y, cond, mask, **kw)
            loss_o = lo.fm_loss(vf_g, kind, x, mask, t, n0, n1, SIGMA)
            loss_o.backward()
            assert torch.equal(loss_r.detach(), loss_o.detach()), (name, kind, loss_r, loss_o)
            tag = kind.replace("-", "").lower()
            out[f"loss_{tag}"] = loss_r.detach().numpy()
            out[f"loss_{tag}_t"] = t.numpy()
            out[f"loss_{tag}_n0"] = n0.numpy()
            if n1 is not None:
                out[f"loss_{tag}_n1"] = n1.numpy()
            gmax = 0.0
            for k, p in m.flows[0].net.named_parameters():
                go = sd_g[k].grad
                gmax = max(gmax, float((p.grad - go).abs().max()))
                assert torch.allclose(p.grad, go, rtol=1e-5, atol=1e-7), (name, kind, k)
            if kind == "FM-OT":      # keep gradients of one loss kind: per-parameter norms always,
                params = list(m.flows[0].net.named_parameters())     # all entries for the small nets
                out["grad_fmot_names"] = np.array([k for k, _ in params])
                out["grad_fmot_norms"] = np.array([float(p.grad.norm()) for _, p in params], dtype="float64")
                if sum(p.numel() for _, p in params) < 100_000:
                    out["grad_fmot"] = torch.cat([p.grad.flatten() for _, p in params]).numpy()
            print(f"  {name:20s} loss[{kind}] = {float(loss_r):.6f}  max|dgrad|={gmax:.2e}")

        # -- sample(): reference plumbing + torchdyn stand-in  vs  oracle integrator
        m = build_reference(ref, ctor, sd, "FM-OT")
        for solver, steps in case["ode"]:
            torch.manual_seed(777)
            s_r = m.sample(B, cond=cond, mask=mask, ode_solver=solver, ode_steps=steps)
            torch.manual_seed(777)
            z = torch.randn(B, N, Fd)
            with torch.no_grad():
                s_o = lo.sample(vf_o, z, mask, solver, steps)
            assert torch.equal(s_r, s_o), (name, solver, (s_r - s_o).abs().max())
            out[f"z_{solver}{steps}"] = z.numpy()
            out[f"sample_{solver}{steps}"] = s_r.numpy()
        meta = dict(ctor=ctor, cfg=cfg.as_dict(), wseed=case["wseed"], weight_norm=wn, cloud=cl,
                    sigma=SIGMA, ode=case["ode"], torch=torch.__version__)
        out["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
        print(f"{name}: ok  ({os.path.getsize(os.path.join(GOLDEN_DIR, name + '.npz'))/1024:.1f} KiB)")

    # zero-multiplicity jet: NaN confined to that jet (epic.py:161,370)
    ctor = CASES["c1_jetnet30"]["ctor"]
    cfg = cfg_from_ctor(ctor)
    sd = eo.synth_state_dict(cfg, 12345)
    x, mask, _ = eo.synth_cloud(3, 30, 3, 55)
    mask[1] = 0
    x = x * mask
    m = build_reference(ref, ctor, sd, "FM-OT")
    with torch.no_grad():
        v = m.flows[0](torch.tensor(0.25), x, mask=mask)
    assert torch.isnan(v[1]).all() and not torch.isnan(v[0]).any() and not torch.isnan(v[2]).any()
    print("zero-multiplicity jet: NaN confined to the jet, ok")


if __name__ == "__main__":
    sys.exit(main())
