"""GPU: the fused training step (loss forward + backward through the C ABI) against torch autograd over the
CPU oracle.  fp32 path: loss relative error <= 1e-5, every parameter gradient relative L2 <= 1e-4."""
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import loss_oracle as lo

from helpers import GOLDEN_CASES, Golden, build_module, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOSS_TOL = 1e-5
GRAD_TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def oracle_loss_and_grads(g, kind, x, mask, cond, t, n0, n1, sigma):
    sd = {k: v.clone().requires_grad_(True) for k, v in g.sd.items()}
    kw = g.oracle_kwargs()
    vf = lambda tt, y: eo.cnf_forward(sd, g.cfg, tt, y, cond, mask, **kw)
    loss = lo.fm_loss(vf, kind, x, mask, t, n0, n1, sigma)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in sd.items()}


def module_grads(m):
    return {k[len("flows.0.net."):]: p.grad.detach().cpu() for k, p in m.named_parameters() if k.startswith("flows.0.net.")}


def run_fused(m, kind, x, mask, cond, t, n0, n1, sigma):
    from particle_fm_b200.training import fm_loss_autograd
    m.zero_grad(set_to_none=True)
    c = None if cond is None else cond.to(DEV)
    loss = fm_loss_autograd(m.flows[0], kind, x.to(DEV), mask.to(DEV), c, t.to(DEV), n0.to(DEV),
                            None if n1 is None else n1.to(DEV), sigma)
    loss.backward()
    return loss.detach().cpu(), module_grads(m)


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("kind", ["FM-OT", "CFM", "droid"])
def test_loss_and_gradients_vs_oracle_autograd(name, kind):
    g = Golden(name)
    m = build_module(g.ctor, g.sd, loss_type=kind, device=DEV)
    gen = torch.Generator().manual_seed(31)
    x = g.x * 5.0 * g.mask
    B = x.shape[0]
    t = torch.rand(B, generator=gen)
    n0 = torch.randn(x.shape, generator=gen)
    n1 = torch.randn(x.shape, generator=gen) if kind == "CFM" else None
    sigma = 1e-4
    ref_loss, ref_g = oracle_loss_and_grads(g, kind, x, g.mask, g.cond, t, n0, n1, sigma)
    loss, got = run_fused(m, kind, x, g.mask, g.cond, t, n0, n1, sigma)
    assert abs(float(loss) - float(ref_loss)) <= LOSS_TOL * abs(float(ref_loss)), (float(loss), float(ref_loss))
    assert set(got) == set(ref_g)
    worst = max((rel_l2(got[k], ref_g[k]), k) for k in ref_g if ref_g[k] is not None and float(ref_g[k].norm()) > 0)
    assert worst[0] < GRAD_TOL, worst
    # gradient of a parameter the loss does not depend on is exactly zero on both sides
    for k, v in ref_g.items():
        if v is not None and float(v.norm()) == 0:
            assert float(got[k].norm()) == 0, k


def test_loss_module_api_and_rng_order():
    """SetFlowMatchingLitModule.training_step: t from the CPU generator, noise on the device;
    the same draws fed to the oracle give the same loss; validation (no_grad) gives the same value, no grads."""
    g = Golden("c1_jetnet30")
    m = build_module(g.ctor, g.sd, loss_type="FM-OT", device=DEV)
    x = (g.x * 5.0 * g.mask).to(DEV)
    mask = g.mask.to(DEV)
    torch.manual_seed(5)
    out = m.training_step((x, mask, torch.zeros(x.shape[0], device=DEV)), 0)
    loss = out["loss"]
    assert loss.requires_grad and loss.dim() == 0
    loss.backward()
    assert all(p.grad is not None for p in m.flows[0].net.parameters())
    torch.manual_seed(5)
    t = torch.rand_like(torch.ones(x.shape[0]))
    n0 = torch.randn_like(x)
    vf = g.oracle_vf()
    ref = lo.fm_loss(vf, "FM-OT", x.cpu(), g.mask, t, n0.cpu(), None, 1e-4)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    torch.manual_seed(5)
    val = m.validation_step((x, mask, torch.zeros(x.shape[0], device=DEV)), 0)["loss"]
    assert not val.requires_grad and abs(float(val) - float(loss)) <= 1e-6 * abs(float(loss))


def test_generic_autograd_of_the_vector_field():
    """CNF.forward under autograd with an arbitrary downstream loss: gradients w.r.t. x and the parameters."""
    g = Golden("cond_lhco_like")
    m = build_module(g.ctor, g.sd, device=DEV)
    gen = torch.Generator().manual_seed(3)
    x = g.x.clone()
    w = torch.randn(x.shape, generator=gen)
    B, N = x.shape[:2]
    t = torch.rand(B, generator=gen).unsqueeze(-1).repeat_interleave(N, dim=1)
    sd = {k: v.clone().requires_grad_(True) for k, v in g.sd.items()}
    xr = x.clone().requires_grad_(True)
    v_ref = eo.cnf_forward(sd, g.cfg, t, xr, g.cond, g.mask, **g.oracle_kwargs())
    (v_ref * w).sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    m.zero_grad(set_to_none=True)
    v = m.flows[0](t.to(DEV), xd, cond=g.cond.to(DEV), mask=g.mask.to(DEV))
    assert rel_l2(v.detach().cpu(), v_ref.detach()) < 1e-5
    (v * w.to(DEV)).sum().backward()
    assert rel_l2(xd.grad.cpu() * g.mask, xr.grad * g.mask) < GRAD_TOL
    assert float((xd.grad.cpu() * (1 - g.mask)).abs().max()) == 0
    got = module_grads(m)
    worst = max((rel_l2(got[k], sd[k].grad), k) for k in sd if float(sd[k].grad.norm()) > 0)
    assert worst[0] < GRAD_TOL, worst


def test_optimizer_step_changes_the_packed_weights():
    """AdamW step -> parameters change in place -> the packed copy is refreshed; loss goes down on a fixed batch."""
    g = Golden("c1_jetnet30")
    m = build_module(g.ctor, g.sd, loss_type="FM-OT", device=DEV)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=5e-5)
    x = (g.x * 5.0 * g.mask).to(DEV)
    mask = g.mask.to(DEV)
    losses = []
    for _ in range(8):
        torch.manual_seed(11)
        opt.zero_grad()
        loss = m.loss(x, mask=mask, cond=None)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses


def test_weight_norm_fold_and_chain_rule_in_the_library():
    """pfm_epic_set_params / pfm_epic_param_grads against torch._weight_norm and its autograd (what the reference runs as
    a forward pre-hook on every linear, epic.py:66-81): folded weights give the same vector field as handing over the
    torch-folded weights, and an arbitrary flat folded-weight gradient maps onto the same d weight_v / d weight_g / d bias."""
    g = Golden("cond_lhco_like")
    m = build_module(g.ctor, g.sd, device=DEV)
    net = m.flows[0].net
    lins = net.linears()
    eng = net.engine(sync_weights=False)
    # forward: library fold vs torch fold
    eng.set_params(lins, key=None)
    t = torch.tensor(0.41, device=DEV)
    with torch.no_grad():
        cond = g.cond.to(DEV)
        code = m.flows[0].time_code(t)
        v_lib = eng.forward(code, g.x.to(DEV), g.mask.to(DEV), cond)
        folded = [lin.folded() for lin in lins]
        eng.set_weights([w for w, _ in folded], [b for _, b in folded], key=None)
        v_torch = eng.forward(code, g.x.to(DEV), g.mask.to(DEV), cond)
    assert rel_l2(v_lib.cpu(), v_torch.cpu()) < 1e-6
    # backward: random flat gradient of the folded weights
    gen = torch.Generator().manual_seed(9)
    flat = torch.randn(eng.grad_size(), generator=gen).to(DEV)
    scale = torch.tensor(0.37, device=DEV)
    got = eng.param_grads(flat, scale, lins)
    views = eng.grad_views(flat)
    for lin, (dv, dg, db), (gw, gb) in zip(lins, got, views):
        if lin.weight_norm:
            v = lin.weight_v.detach().clone().requires_grad_(True)
            gg = lin.weight_g.detach().clone().requires_grad_(True)
            w = torch._weight_norm(v, gg, 0)
            (w * gw.view_as(w)).sum().mul(scale).backward()
            assert rel_l2(dv.cpu(), v.grad.cpu()) < 1e-5
            assert rel_l2(dg.cpu().flatten(), gg.grad.cpu().flatten()) < 1e-5
        else:
            assert dg is None and rel_l2(dv.cpu(), (gw.view_as(dv) * scale).cpu()) < 1e-6
        assert rel_l2(db.cpu(), (gb * scale).cpu()) < 1e-6


def test_stale_backward_after_inference_call_raises():
    """an inference forward / sample() between forward_train and backward rewrites the handle's plan buffers;
    the pending backward must fail (ticket + PFM_ERR_STATE) instead of running on a foreign plan."""
    g = Golden("c1_jetnet30")
    m = build_module(g.ctor, g.sd, device=DEV)
    B, N = g.x.shape[:2]
    t = torch.rand(B, generator=torch.Generator().manual_seed(3)).unsqueeze(-1).repeat_interleave(N, dim=1)
    xd = g.x.to(DEV).requires_grad_(True)
    v = m.flows[0](t.to(DEV), xd, cond=None, mask=g.mask.to(DEV))            # generic autograd path: forward_train
    other_mask = torch.ones_like(g.mask)
    other_mask[:, 7:] = 0
    with torch.no_grad():                                                     # inference call with a different mask
        m.flows[0](torch.tensor(0.3, device=DEV), g.x.to(DEV) * other_mask.to(DEV), cond=None, mask=other_mask.to(DEV))
    with pytest.raises(RuntimeError, match="overwritten|forward_train|PFM"):
        v.sum().backward()
    # and directly at the C ABI: the handle itself refuses (train_B was reset by the inference call)
    eng = m.flows[0].net.engine()
    code = m.flows[0].time_code(t[:, 0])
    out, ticket, saved = eng.forward_train(code.to(DEV), g.x.to(DEV), g.mask.to(DEV), None)
    eng.forward(code[:1].to(DEV), g.x.to(DEV), other_mask.to(DEV), None)
    eng._ticket = ticket                                                      # defeat the Python-side guard
    with pytest.raises(Exception, match="no matching pfm_epic_forward_train"):
        eng.backward(ticket, saved, torch.ones_like(out), True, True)


def test_in_place_data_updates_reach_the_sampler():
    """p.data.mul_() bumps neither _version nor data_ptr; decode() re-syncs unconditionally and
    invalidate_weights() covers the per-step forward."""
    g = Golden("c1_jetnet30")
    m = build_module(g.ctor, g.sd, device=DEV)
    z = (g.t("z_euler100") * g.mask).to(DEV)
    mask = g.mask.to(DEV)
    a = m.flows[0].decode(z, None, mask, "euler", 5)
    with torch.no_grad():
        for p in m.flows[0].net.fc_l3.parameters():
            p.data.mul_(0.5)
    b = m.flows[0].decode(z, None, mask, "euler", 5)
    assert not torch.equal(a, b), "sample() used stale packed weights after an in-place .data update"
    sd2 = {k[len("flows.0.net."):]: v.detach().cpu() for k, v in m.state_dict().items() if k.startswith("flows.0.net.")}
    ref = lo.sample(g.oracle_vf(sd=sd2), z.cpu(), g.mask, "euler", 5)
    assert rel_l2(b.cpu(), ref) < 1e-4
    t = torch.tensor(0.5, device=DEV)
    with torch.no_grad():
        v1 = m.flows[0](t, z, cond=None, mask=mask)
        for p in m.flows[0].net.fc_l3.parameters():
            p.data.mul_(2.0)
        m.flows[0].net.invalidate_weights()
        v2 = m.flows[0](t, z, cond=None, mask=mask)
    assert not torch.equal(v1, v2)


def test_training_step_loss_per_jettype_branch():
    """flow_matching_module.py:526-551: with datamodule.hparams.loss_per_jettype the step logs one extra loss per jet type
    on the slice of that type, consuming the RNG streams like the reference's extra self.loss(...) calls."""
    import types
    g = Golden("cond_jetclass_like")
    m = build_module(g.ctor, g.sd, loss_type="FM-OT", device=DEV)
    B = g.x.shape[0]
    C = g.cond.shape[1]
    cond = g.cond.clone()
    cond[:, 0] = (torch.arange(B) % 2 == 0).float()                          # one-hot jet-type labels in columns 0 / 1
    cond[:, 1] = 1 - cond[:, 0]
    names = ["jet_type_label_A", "jet_type_label_B"] + [f"c{i}" for i in range(C - 2)]
    dm = types.SimpleNamespace(hparams=types.SimpleNamespace(variable_jet_sizes=True, loss_per_jettype=True,
                                                             used_jet_types=["A", "B"]), names_conditioning=names)
    m.trainer = types.SimpleNamespace(datamodule=dm)
    logged = {}
    m.log = lambda name, value, **kw: logged.__setitem__(name, value)
    x, mask = (g.x * 5.0 * g.mask).to(DEV), g.mask.to(DEV)
    torch.manual_seed(21)
    out = m.training_step((x, mask, cond.to(DEV)), 0)
    assert set(logged) == {"train/loss", "train/loss_A", "train/loss_B"}
    # the same sequence of draws replayed against the oracle
    torch.manual_seed(21)
    refs = {}
    for name, sel in (("train/loss", torch.ones(B, dtype=torch.bool)), ("train/loss_A", cond[:, 0] == 1), ("train/loss_B", cond[:, 1] == 1)):
        xs = x[sel.to(DEV)]
        t = torch.rand_like(torch.ones(xs.shape[0]))
        n0 = torch.randn_like(xs)
        vf = g.oracle_vf(cond=cond[sel], mask=g.mask[sel])
        refs[name] = float(lo.fm_loss(vf, "FM-OT", xs.cpu(), g.mask[sel], t, n0.cpu(), None, 1e-4))
    for k, r in refs.items():
        assert abs(float(logged[k]) - r) <= 1e-5 * abs(r), (k, float(logged[k]), r)
    assert out["loss"].requires_grad and not logged["train/loss_A"].requires_grad
    m.current_epoch = 3                                                       # not a multiple of 20: branch off
    logged.clear()
    m.training_step((x, mask, cond.to(DEV)), 0)
    assert set(logged) == {"train/loss"}
