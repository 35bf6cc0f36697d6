"""Droid set transformers.
CPU: the oracle against the golden vectors recorded from the reference; the host mirror's state_dict layout.
GPU: the CUDA path (C ABI pfm_tf_*) against the golden vectors and the oracle.  fp32 tolerance: 2e-5 relative L2 per
evaluation over the REAL particles (padding is skipped; the reference leaves unmasked values in padded slots that its
callers multiply away)."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from oracle import droid_oracle as do
from oracle import epic_oracle as eo
from oracle import ode_oracle as oo

from helpers import GOLDEN_DIR, rel_l2

DEV = "cuda:0"
CASES = ["droid_full_n30", "droid_full_n150_cond", "droid_cross_n30", "droid_cross_n150_cond"]
STEP_TOL = 2e-5

NET_CONFIG = {     # configs/model/fm_droid_transformer.yaml:15-34, fm_droid_crossattention.yaml:15-35
    "full": dict(node_embd_config=dict(act_h="lrlu", nrm="layer"), ctxt_embd_config=dict(outp_dim=64, act_h="lrlu", nrm="layer"),
                 te_config=dict(model_dim=256, num_layers=3, mha_config=dict(num_heads=16, init_zeros=True, do_layer_norm=True),
                                dense_config=dict(act_h="lrlu", nrm="layer", output_init_zeros=True)),
                 outp_embd_config=dict(act_h="lrlu", nrm="layer", output_init_zeros=True)),
    "cross": dict(node_embd_config=dict(act_h="lrlu", nrm="layer"), ctxt_embd_config=dict(outp_dim=64, act_h="lrlu", nrm="layer"),
                  cae_config=dict(model_dim=128, num_layers=8, mha_config=dict(num_heads=16, init_zeros=True, do_layer_norm=True),
                                  dense_config=dict(hddn_dim=256, act_h="lrlu", nrm="layer", output_init_zeros=True)),
                  outp_embd_config=dict(act_h="lrlu", nrm="layer", output_init_zeros=True)),
}
MODEL = {"full": "droid_fulltransformer", "cross": "droid_fullcrossattention"}


class G:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.meta = json.loads(str(z["meta"]))
        self.a = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta" and z[k].dtype.kind in "fiu"}
        self.cfg = do.DroidCfg(**self.meta["cfg"])
        self.sd = do.synth_state_dict(self.cfg, self.meta["wseed"])
        self.kind, self.N = self.meta["kind"], self.meta["N"]
        self.cond = self.a.get("cond")


def build(g, device=None, N=None):
    from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
    m = SetFlowMatchingLitModule(optimizer=None, model=MODEL[g.kind], features=3, num_particles=N or g.N, frequencies=16,
                                 t_emb="cosine", add_time_to_input=True, global_cond_dim=g.cfg.cond_dim,
                                 net_config=copy.deepcopy(NET_CONFIG[g.kind]))
    m.flows[0].net.load_state_dict(g.sd, strict=True)
    return m.to(device) if device else m


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_golden(name):
    g = G(name)
    x, mask = g.a["x"], g.a["mask"]
    tb = g.a["t_train"].unsqueeze(-1).repeat_interleave(g.N, dim=1)
    with torch.no_grad():
        v_s = do.cnf_forward(g.sd, g.cfg, g.a["t_sample"], x, g.cond, mask)
        v_t = do.cnf_forward(g.sd, g.cfg, tb, x, g.cond, mask)
        end = oo.integrate(lambda t, y: do.cnf_forward(g.sd, g.cfg, t, y, g.cond, mask), x, 5, "euler")
    assert rel_l2(v_s, g.a["v_sample"]) < 1e-6 and rel_l2(v_t, g.a["v_train"]) < 1e-6
    assert rel_l2(end, g.a["sample_euler5"]) < 1e-5
    # only keys are masked: real tokens do not depend on what sits in the padded slots
    junk = x + (1 - mask) * 7.0
    with torch.no_grad():
        v_j = do.cnf_forward(g.sd, g.cfg, g.a["t_sample"], junk, g.cond, mask)
    assert rel_l2(v_j * mask, v_s * mask) < 1e-6


@pytest.mark.parametrize("kind", ["full", "cross"])
def test_host_mirror_state_dict_layout(kind):
    """Same parameter names, shapes and ORDER as the reference's CNF.net (the order is the C ABI's canonical order)."""
    from particle_fm_b200.models.flow_matching_module import CNF
    cfg = do.yaml_cfg(kind, 3, 5)
    cnf = CNF(model=MODEL[kind], features=3, num_particles=30, frequencies=16, t_emb="cosine", add_time_to_input=True,
              global_cond_dim=5, net_config=copy.deepcopy(NET_CONFIG[kind]))
    got = [(n, tuple(p.shape)) for n, p in cnf.net.named_parameters()]
    assert got == do.param_shapes(cfg)
    cnf.net.load_state_dict(do.synth_state_dict(cfg, 3), strict=True)
    # the zero-initialised layers of the YAML (init_zeros / output_init_zeros) start at zero like the reference's
    fresh = CNF(model=MODEL[kind], features=3, num_particles=30, frequencies=16, t_emb="cosine", add_time_to_input=True,
                net_config=copy.deepcopy(NET_CONFIG[kind]))
    assert float(fresh.net.outp_embd.output_block.block[0].weight.abs().max()) == 0
    with pytest.raises(NotImplementedError):
        CNF(model=MODEL[kind], features=3, frequencies=16, t_emb="cosine",
            net_config={**copy.deepcopy(NET_CONFIG[kind]), "node_embd_config": dict(act_h="relu", nrm="layer")})


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_vector_field_and_sample_vs_golden(name, lib_built):
    g = G(name)
    m = build(g, DEV)
    cnf = m.flows[0]
    x, mask = g.a["x"], g.a["mask"]
    cd = None if g.cond is None else g.cond.to(DEV)
    tb = g.a["t_train"].unsqueeze(-1).repeat_interleave(g.N, dim=1)
    with torch.no_grad():
        v_s = cnf(g.a["t_sample"].to(DEV), x.to(DEV), cond=cd, mask=mask.to(DEV)).cpu()
        v_t = cnf(tb.to(DEV), x.to(DEV), cond=cd, mask=mask.to(DEV)).cpu()
        end = cnf.decode(x.to(DEV), cd, mask.to(DEV), "euler", 5).cpu()
    es, et = rel_l2(v_s, g.a["v_sample"] * mask), rel_l2(v_t, g.a["v_train"] * mask)
    ee = rel_l2(end * mask, g.a["sample_euler5"] * mask)
    print(f"{name}: sampling-mode {es:.2e} training-mode {et:.2e} euler-5 end point {ee:.2e}")
    assert es < STEP_TOL and et < STEP_TOL and ee < 1e-4
    assert (v_s * (1 - mask)).abs().max() == 0          # padded slots come back as 0
    # int64 mask, junk in the padded slots
    with torch.no_grad():
        v_j = cnf(g.a["t_sample"].to(DEV), (x + (1 - mask) * 5.0).to(DEV), cond=cd, mask=mask.long().to(DEV)).cpu()
    assert torch.equal(v_j, v_s)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["droid_full_n30", "droid_cross_n150_cond"])
def test_cuda_many_jets_midpoint_and_properties(name, lib_built):
    g = G(name)
    N, B = g.N, 70
    m = build(g, DEV)
    x, mask, cond = eo.synth_cloud(B, N, 3, 4242, cond_dim=g.cfg.cond_dim)
    cd = None if cond is None else cond.to(DEV)
    t = torch.tensor(0.61)
    f = lambda xx, mm, cc: m.flows[0](t.to(DEV), xx.to(DEV), cond=cc, mask=mm.to(DEV)).cpu()
    with torch.no_grad():
        v = f(x, mask, cd)
        sub = torch.arange(0, B, 9)
        vo = do.cnf_forward(g.sd, g.cfg, t, x[sub], None if cond is None else cond[sub], mask[sub])
        assert rel_l2(v[sub], vo * mask[sub]) < STEP_TOL
        pb = torch.randperm(B, generator=torch.Generator().manual_seed(1))
        assert rel_l2(f(x[pb], mask[pb], None if cd is None else cd[pb.to(DEV)]), v[pb]) < 2e-6      # batch order
        pn = torch.randperm(N, generator=torch.Generator().manual_seed(2))
        assert rel_l2(f(x[:, pn], mask[:, pn], cd), v[:, pn]) < 5e-6                                   # permutation equivariance
        # whole midpoint integration through sample(): same CPU-generator noise as the oracle run
        torch.manual_seed(11)
        s = m.sample(8, cond=None if cond is None else cond[:8], mask=mask[:8], ode_solver="midpoint", ode_steps=4).cpu()
        torch.manual_seed(11)
        z = torch.randn(8, N, 3) * mask[:8]
        ref = oo.integrate(lambda tt, y: do.cnf_forward(g.sd, g.cfg, tt, y, None if cond is None else cond[:8], mask[:8]) * mask[:8],
                           z, 4, "midpoint")
    assert rel_l2(s, ref * mask[:8]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["droid_full_n150_cond", "droid_cross_n150_cond"])
def test_cuda_bf16_tensor_core_linears(name, lib_built):
    """PFM_PREC_BF16: the per-token linears on tcgen05 (bf16 operands, fp32 accumulate); tolerance 2e-2 per evaluation."""
    g = G(name)
    N, B = g.N, 40                                   # enough rows (> 256) for the tensor-core path to engage
    m = build(g, DEV)
    x, mask, cond = eo.synth_cloud(B, N, 3, 777, cond_dim=g.cfg.cond_dim)
    cd = None if cond is None else cond.to(DEV)
    t = torch.tensor(0.43)
    with torch.no_grad():
        v32 = m.flows[0](t.to(DEV), x.to(DEV), cond=cd, mask=mask.to(DEV)).cpu()
        l32 = m.flows[0].net.engine().last_launches()
        m.set_precision("bf16")
        v16 = m.flows[0](t.to(DEV), x.to(DEV), cond=cd, mask=mask.to(DEV)).cpu()
        sub = torch.arange(0, B, 7)
        vo = do.cnf_forward(g.sd, g.cfg, t, x[sub], None if cond is None else cond[sub], mask[sub]) * mask[sub]
        s16 = m.flows[0].decode(x.to(DEV), cd, mask.to(DEV), "midpoint", 4).cpu()
        m.set_precision("fp32")
        s32 = m.flows[0].decode(x.to(DEV), cd, mask.to(DEV), "midpoint", 4).cpu()
    e = rel_l2(v16[sub], vo)
    print(f"{name}: bf16 vector field rel-L2 {e:.2e} (fp32 path {rel_l2(v32[sub], vo):.2e}); midpoint-4 end point vs fp32 {rel_l2(s16, s32):.2e}")
    assert 1e-5 < e < 2e-2                            # really a different (bf16) arithmetic, within tolerance
    assert rel_l2(s16, s32) < 5e-3                  # 3 big midpoint steps: the end point carries the per-evaluation error
    assert (v16 * (1 - mask)).abs().max() == 0 and l32 > 0
