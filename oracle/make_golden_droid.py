"""Oracle (test infrastructure): golden vectors of the droid set transformers from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_droid
Builds the reference's own CNF(model="droid_fulltransformer" | "droid_fullcrossattention") with the net_config of
configs/model/fm_droid_transformer.yaml / fm_droid_crossattention.yaml (through oracle/ref_shim.py), loads
deterministic synthetic weights (strict), records the vector field in sampling mode (0-dim t) and training mode
(per-jet t) plus a short Euler integration, and asserts that oracle/droid_oracle.py reproduces them.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import droid_oracle as do
from . import epic_oracle as eo
from . import ode_oracle as oo
from . import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

NET_CONFIG = {
    "full": dict(node_embd_config=dict(act_h="lrlu", nrm="layer"),
                 ctxt_embd_config=dict(outp_dim=64, act_h="lrlu", nrm="layer"),
                 te_config=dict(model_dim=256, num_layers=3,
                                mha_config=dict(num_heads=16, init_zeros=True, do_layer_norm=True),
                                dense_config=dict(act_h="lrlu", nrm="layer", output_init_zeros=True)),
                 outp_embd_config=dict(act_h="lrlu", nrm="layer", output_init_zeros=True)),
    "cross": dict(node_embd_config=dict(act_h="lrlu", nrm="layer"),
                  ctxt_embd_config=dict(outp_dim=64, act_h="lrlu", nrm="layer"),
                  cae_config=dict(model_dim=128, num_layers=8,
                                  mha_config=dict(num_heads=16, init_zeros=True, do_layer_norm=True),
                                  dense_config=dict(hddn_dim=256, act_h="lrlu", nrm="layer", output_init_zeros=True)),
                  outp_embd_config=dict(act_h="lrlu", nrm="layer", output_init_zeros=True)),
}
MODEL_NAME = {"full": "droid_fulltransformer", "cross": "droid_fullcrossattention"}

CASES = {
    "droid_full_n30": dict(kind="full", N=30, B=4, cond_dim=0, seed=201, wseed=31),
    "droid_full_n150_cond": dict(kind="full", N=150, B=2, cond_dim=5, seed=202, wseed=32),     # LHCO jets_transformer: cond 5
    "droid_cross_n30": dict(kind="cross", N=30, B=4, cond_dim=0, seed=203, wseed=33),
    "droid_cross_n150_cond": dict(kind="cross", N=150, B=2, cond_dim=5, seed=204, wseed=34),
}


def main():
    import copy
    R = ref_shim.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, c in CASES.items():
        cfg = do.yaml_cfg(c["kind"], 3, c["cond_dim"])
        sd = do.synth_state_dict(cfg, c["wseed"])
        cnf = R.fm.CNF(model=MODEL_NAME[c["kind"]], features=3, num_particles=c["N"], frequencies=16, t_emb="cosine",
                       add_time_to_input=True, global_cond_dim=c["cond_dim"], net_config=copy.deepcopy(NET_CONFIG[c["kind"]]))
        missing = cnf.net.load_state_dict(sd, strict=True)
        assert set(dict(cnf.net.named_parameters())) == set(sd), "parameter inventory differs from the reference"
        cnf.eval()
        x, mask, cond = eo.synth_cloud(c["B"], c["N"], 3, c["seed"], cond_dim=c["cond_dim"])
        rc = cond if cond is not None else torch.zeros(c["B"], 0)        # cond=None crashes the reference
        g = torch.Generator().manual_seed(c["seed"])
        t_s = torch.rand((), generator=g)
        t_b = torch.rand(c["B"], generator=g)
        t_bn = t_b.unsqueeze(-1).repeat_interleave(c["N"], dim=1)
        with torch.no_grad():
            v_s = cnf(t_s, x, cond=rc, mask=mask)
            v_t = cnf(t_bn, x, cond=rc, mask=mask)
            o_s = do.cnf_forward(sd, cfg, t_s, x, cond, mask)
            o_t = do.cnf_forward(sd, cfg, t_bn, x, cond, mask)
            end = oo.integrate(lambda t, y: cnf(t, y, cond=rc, mask=mask), x, 5, "euler")
        for a, b, what in ((o_s, v_s, "sampling"), (o_t, v_t, "training")):
            err = float((a - b).abs().max())
            assert err <= 1e-6 * max(1.0, float(b.abs().max())), (name, what, err)
        # -- training: the reference's own loss modules + autograd vs autograd through the oracle restatement, same draws
        # (losses.py sums (v - u)^2 over EVERY slot and the droid nets do not mask their output, so padded slots count)
        import torch.nn as nn
        from . import loss_oracle as lo
        extra = {}
        kw = dict(t_emb="cosine", frequencies=16)
        for kind, cls in (("FM-OT", "FlowMatchingLoss"), ("CFM", "ConditionalFlowMatchingLoss"), ("droid", "DroidLoss")):
            loss_mod = getattr(R.losses, cls)(flows=nn.ModuleList([cnf]), sigma=1e-4)
            cnf.zero_grad()
            torch.manual_seed(4242)
            loss_r = loss_mod(x, mask=mask, cond=rc)
            loss_r.backward()
            torch.manual_seed(4242)
            t, n0, n1 = lo.draw_loss_randoms(kind, x)
            sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            loss_o = lo.fm_loss(lambda tt, y: do.cnf_forward(sd_g, cfg, tt, y, cond, mask, **kw), kind, x, mask, t, n0, n1, 1e-4)
            loss_o.backward()
            assert abs(float(loss_r) - float(loss_o)) <= 1e-6 * abs(float(loss_r)), (name, kind, loss_r, loss_o)
            gmax, gref = 0.0, 0.0
            params = list(cnf.net.named_parameters())
            for k, prm in params:
                gmax = max(gmax, float((prm.grad - sd_g[k].grad).abs().max()))
                gref = max(gref, float(prm.grad.abs().max()))
            assert gmax <= 2e-5 * gref, (name, kind, gmax, gref)
            tag = kind.replace("-", "").lower()
            extra[f"loss_{tag}"] = loss_r.detach().numpy()
            extra[f"loss_{tag}_t"] = t.numpy()
            if kind == "FM-OT":       # the draws and a digest of the reference's gradient for one loss kind
                extra["loss_n0"] = n0.numpy()
                extra["grad_fmot_names"] = np.array([k for k, _ in params])
                extra["grad_fmot_norms"] = np.array([float(prm.grad.double().norm()) for _, prm in params], dtype="float64")
                extra["grad_fmot_sums"] = np.array([float(prm.grad.double().sum()) for _, prm in params], dtype="float64")
                extra["grad_fmot_vectors"] = torch.cat([prm.grad.flatten() for _, prm in params if prm.dim() == 1]).numpy()
            print(f"  {name:22s} loss[{kind}] = {float(loss_r):.6f}  max|grad diff| = {gmax:.2e} (max|grad| {gref:.2e})")
        out = dict(x=x.numpy(), mask=mask.numpy(), t_sample=t_s.numpy(), t_train=t_b.numpy(), v_sample=v_s.numpy(),
                   v_train=v_t.numpy(), sample_euler5=end.numpy(),
                   meta=np.array(json.dumps(dict(cfg=cfg.as_dict(), wseed=c["wseed"], N=c["N"], kind=c["kind"]))))
        if cond is not None:
            out["cond"] = cond.numpy()
        out.update(extra)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
        print(f"{name}: oracle == reference (max |v| {float(v_s.abs().max()):.3f}), saved")


if __name__ == "__main__":
    main()
