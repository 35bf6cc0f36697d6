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


def run_fused(m, kind, x, mask, cond, t, n0, n1, sigma, mode="cuda_cores"):
    from particle_fm_b200.training import fm_loss_autograd
    m.zero_grad(set_to_none=True)
    m.flows[0].net.engine().set_train_mode(mode)
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
    if g.cfg.hid != 128:
        return
    # hid == 128: the default training path is the tensor-core program (csrc/epic_train_tc.cu).  Same gates -- unless the two
    # CUDA evaluations of the forward put a pre-activation on different sides of leaky_relu's kink (|z| within fp32 rounding
    # of 0: a measure-zero discontinuity of the derivative, not an arithmetic error); that is detected from the saved
    # activations of both paths and then the draw is repeated with the next seed.
    rows = int(g.mask.sum())
    N = x.shape[1]
    for seed in (31, 32, 33, 34):
        gen = torch.Generator().manual_seed(seed)
        t = torch.rand(B, generator=gen)
        n0 = torch.randn(x.shape, generator=gen)
        n1 = torch.randn(x.shape, generator=gen) if kind == "CFM" else None
        run_fused(m, kind, x, g.mask, g.cond, t, n0, n1, sigma, mode="cuda_cores")
        a_cc = _act_arrays(m.flows[0].net.engine(), B, N, rows)
        loss, got = run_fused(m, kind, x, g.mask, g.cond, t, n0, n1, sigma, mode="auto")
        a_tc = _act_arrays(m.flows[0].net.engine(), B, N, rows)
        flips = sum(int(((u > 0) != (v > 0)).sum()) for u, v in zip(a_cc, a_tc))
        if flips:
            continue
        ref_loss, ref_g = oracle_loss_and_grads(g, kind, x, g.mask, g.cond, t, n0, n1, sigma)
        assert abs(float(loss) - float(ref_loss)) <= LOSS_TOL * abs(float(ref_loss)), (float(loss), float(ref_loss))
        worst = max((rel_l2(got[k], ref_g[k]), k) for k in ref_g if ref_g[k] is not None and float(ref_g[k].norm()) > 0)
        assert worst[0] < GRAD_TOL, (seed, worst)
        break
    else:
        raise AssertionError("every draw had a kink-ambiguous pre-activation")


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


def _big_batch(kind, seed=77, B=96, N=150):
    gen = torch.Generator().manual_seed(seed)
    n_real = torch.randint(1, N + 1, (B,), generator=gen)
    n_real[3] = N; n_real[5] = 1
    mask = (torch.arange(N)[None, :] < n_real[:, None]).float().unsqueeze(-1)
    x = torch.randn(B, N, 3, generator=gen) * 5.0 * mask
    t = torch.rand(B, generator=gen)
    n0 = torch.randn(B, N, 3, generator=gen)
    n1 = torch.randn(B, N, 3, generator=gen) if kind == "CFM" else None
    return x, mask, t, n0, n1


def _act_arrays(eng, B, N, rows):
    """saved post-activations [stage][row][128] of the last training forward (valid rows only)"""
    H, L = eng.dims.hid, eng.dims.layers
    SS = B * N * H
    a = eng.debug_array("act", (2 + 2 * L) * SS)
    return [a[s * SS:s * SS + rows * H] for s in range(2 + 2 * L)]


@pytest.mark.parametrize("kind", ["FM-OT", "CFM"])
def test_tensor_core_training_forward_matches_cuda_core_forward(kind):
    """hid == 128: the training forward runs as tcgen05 GEMM passes over the packed particles (csrc/epic_train_tc.cu, 3-term
    bf16 split).  Every saved activation agrees with the fused fp32 CUDA-core kernel to 1e-5 (relative L2 per stage) and the
    loss to 1e-6, on a batch spanning many 128-row tiles with ragged masks (a full jet and a 1-particle jet included)."""
    from particle_fm_b200.training import fm_loss_autograd
    g = Golden("c2_jetnet150")
    x, mask, t, n0, n1 = _big_batch(kind)
    B, N = x.shape[0], x.shape[1]
    rows = int(mask.sum())
    res = {}
    for mode in ("cuda_cores", "auto"):
        m = build_module(g.ctor, g.sd, loss_type=kind, device=DEV)
        eng = m.flows[0].net.engine()
        eng.set_train_mode(mode)
        with torch.no_grad():
            loss = fm_loss_autograd(m.flows[0], kind, x.to(DEV), mask.to(DEV), None, t.to(DEV), n0.to(DEV),
                                    None if n1 is None else n1.to(DEV), 1e-4)
        res[mode] = (float(loss), _act_arrays(eng, B, N, rows))
    (l0, a0), (l1, a1) = res["cuda_cores"], res["auto"]
    assert abs(l1 - l0) <= 1e-6 * abs(l0), (l0, l1)
    for s_, (u, v) in enumerate(zip(a0, a1)):
        assert rel_l2(torch.from_numpy(v), torch.from_numpy(u)) < 1e-5, s_


def test_tensor_core_training_backward_matches_cuda_core_backward():
    """Same saved forward, two backward implementations: the tcgen05 program and the fused fp32 CUDA-core kernel agree to
    1e-5 on dL/dx and 1e-4 on every weight gradient.  (Sharing the forward removes the one legitimate source of larger
    differences between two fp32 evaluation orders: a pre-activation within rounding of 0 taking the other branch of
    leaky_relu'.)"""
    g = Golden("c2_jetnet150")
    x, mask, t, n0, _ = _big_batch("FM-OT", seed=78, B=64)
    B, N = x.shape[0], x.shape[1]
    m = build_module(g.ctor, g.sd, device=DEV)
    cnf = m.flows[0]
    eng = cnf.net.engine()
    code = cnf.time_code(t.to(DEV))
    gen = torch.Generator().manual_seed(5)
    gout = torch.randn(B, N, 3, generator=gen).to(DEV) * mask.to(DEV)
    out = {}
    for mode in ("cuda_cores", "auto"):
        eng.set_train_mode("cuda_cores")
        _, ticket, saved = eng.forward_train(code, x.to(DEV), mask.to(DEV), None)
        eng.set_train_mode(mode)
        gx, flat = eng.backward(ticket, saved, gout, True, True)
        out[mode] = (gx.cpu(), flat.cpu())
    eng.set_train_mode("auto")
    assert rel_l2(out["auto"][0], out["cuda_cores"][0]) < 1e-5
    off = 0
    for k, (o, i) in enumerate(eng.linear_shapes()):
        n = o * i + o
        a, b = out["auto"][1][off:off + n], out["cuda_cores"][1][off:off + n]
        if float(b.norm()) > 0:
            assert rel_l2(a, b) < GRAD_TOL, k
        off += n
    # and the other way round: a tensor-core forward differentiated by both backward paths
    eng.set_train_mode("auto")
    _, ticket, saved = eng.forward_train(code, x.to(DEV), mask.to(DEV), None)
    gx_a, flat_a = eng.backward(ticket, saved, gout, True, True)
    _, ticket, saved = eng.forward_train(code, x.to(DEV), mask.to(DEV), None)
    eng.set_train_mode("cuda_cores")
    gx_b, flat_b = eng.backward(ticket, saved, gout, True, True)
    eng.set_train_mode("auto")
    assert rel_l2(gx_a, gx_b) < 1e-5 and rel_l2(flat_a, flat_b) < 1e-5


@pytest.mark.parametrize("case", ["one_full_jet", "tiny_rows", "jets_longer_than_a_tile", "no_mask", "conditioned"])
def test_tensor_core_training_edge_shapes(case):
    """Shapes at the edges of the tensor-core training program's tiling (128 packed particles per tile): a single jet, fewer
    rows than one tile, jets spanning three tiles (N = 279), mask=None, global + local conditioning.  Loss and dL/dx against the
    fused fp32 CUDA-core kernels; weight gradients on a shared forward (see the test above)."""
    g = Golden("c2_jetnet150")
    gen = torch.Generator().manual_seed(4242)
    B, N, cg, cl = {"one_full_jet": (1, 150, 0, 0), "tiny_rows": (3, 30, 0, 0), "jets_longer_than_a_tile": (5, 279, 0, 0),
                    "no_mask": (4, 150, 0, 0), "conditioned": (9, 150, 2, 2)}[case]
    n_real = torch.randint(1, N + 1, (B,), generator=gen)
    if case == "tiny_rows":
        n_real = torch.tensor([1, 2, 30])
    if case == "jets_longer_than_a_tile":
        n_real[0] = 279; n_real[1] = 129; n_real[2] = 128
    if case in ("one_full_jet", "no_mask"):
        n_real[:] = N
    mask = (torch.arange(N)[None, :] < n_real[:, None]).float().unsqueeze(-1)
    ctor = dict(g.ctor, num_particles=N)
    sd = g.sd
    if cg or cl:
        ctor.update(global_cond_dim=cg, local_cond_dim=cl)
        sd = eo.synth_state_dict(eo.EpicCfg(**{**g.meta["cfg"], "global_cond_dim": cg, "local_cond_dim": cl}), 77)
    m = build_module(ctor, sd, device=DEV)
    cnf = m.flows[0]
    eng = cnf.net.engine()
    x = (torch.randn(B, N, 3, generator=gen) * 5.0 * mask).to(DEV)
    t = torch.rand(B, generator=gen).to(DEV)
    cond = torch.randn(B, max(cg, cl), generator=gen).to(DEV) if (cg or cl) else None
    code = cnf.time_code(t)
    gout = (torch.randn(B, N, 3, generator=gen) * mask).to(DEV)
    mk = None if case == "no_mask" else mask.to(DEV)
    outs = {}
    for mode in ("cuda_cores", "auto"):
        eng.set_train_mode(mode)
        out, ticket, saved = eng.forward_train(code, x, mk, cond)
        outs[mode] = out.cpu()
    assert torch.isfinite(outs["auto"]).all()
    assert rel_l2(outs["auto"], outs["cuda_cores"]) < 1e-5
    grads = {}
    for mode in ("cuda_cores", "auto"):
        eng.set_train_mode("auto")
        _, ticket, saved = eng.forward_train(code, x, mk, cond)
        eng.set_train_mode(mode)
        gx, flat = eng.backward(ticket, saved, gout, True, True)
        grads[mode] = (gx.cpu(), flat.cpu())
    eng.set_train_mode("auto")
    assert rel_l2(grads["auto"][0], grads["cuda_cores"][0]) < 1e-5
    assert rel_l2(grads["auto"][1], grads["cuda_cores"][1]) < 1e-5


@pytest.mark.parametrize("mode", ["cuda_cores", "auto"])
def test_jet_without_particles_poisons_the_loss_like_the_reference(mode):
    """A jet whose mask is all zero: the reference's pooled mean is 0/0, its vector field for that jet is NaN and so is the
    batch loss (NaN * mask stays NaN).  Both training paths reproduce that instead of silently skipping the jet; the other
    jets of the plain forward are unaffected."""
    from particle_fm_b200.training import fm_loss_autograd
    g = Golden("c2_jetnet150")
    gen = torch.Generator().manual_seed(1)
    B, N = 6, 150
    n_real = torch.tensor([10, 0, 150, 3, 0, 77])
    mask = (torch.arange(N)[None, :] < n_real[:, None]).float().unsqueeze(-1)
    x = torch.randn(B, N, 3, generator=gen) * 5 * mask
    t = torch.rand(B, generator=gen)
    n0 = torch.randn(B, N, 3, generator=gen)
    vf = lambda tt, y: eo.cnf_forward(g.sd, g.cfg, tt, y, None, mask, **g.oracle_kwargs())
    with torch.no_grad():
        assert torch.isnan(lo.fm_loss(vf, "FM-OT", x, mask, t, n0, None, 1e-4))
        want = vf(t[:, None].expand(B, N), x)            # per-jet times, broadcast over the particles as the loss does
    m = build_module(g.ctor, g.sd, device=DEV)
    cnf = m.flows[0]
    eng = cnf.net.engine()
    eng.set_train_mode(mode)
    with torch.no_grad():
        loss = fm_loss_autograd(cnf, "FM-OT", x.to(DEV), mask.to(DEV), None, t.to(DEV), n0.to(DEV), None, 1e-4)
    assert torch.isnan(loss)
    out, _, _ = eng.forward_train(cnf.time_code(t.to(DEV)), x.to(DEV), mask.to(DEV), None)
    out = out.cpu()
    keep = [0, 2, 3, 5]
    assert torch.isnan(out[[1, 4]]).all() and torch.isnan(want[[1, 4]]).all()
    assert rel_l2(out[keep], want[keep]) < 1e-5
    eng.set_train_mode("auto")


def test_fused_clip_adamw_matches_torch():
    """particle_fm_b200.optim.FusedClipAdamW (pfm_clip_adamw: global-norm clip + AdamW + EMA over flat buffers, two launches)
    against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW + the reference's EMA recurrence (callbacks/ema.py:77-81) on
    the real training step of the default net, three steps."""
    from particle_fm_b200.optim import FusedClipAdamW
    g = Golden("c1_jetnet30")
    x = (g.x * 5.0 * g.mask).to(DEV)
    mask = g.mask.to(DEV)
    ma = build_module(g.ctor, g.sd, device=DEV)
    mb = build_module(g.ctor, g.sd, device=DEV)
    oa = torch.optim.AdamW(ma.parameters(), lr=1e-3, weight_decay=5e-5)
    ob = FusedClipAdamW(mb.parameters(), lr=1e-3, weight_decay=5e-5, max_grad_norm=0.5, ema_decay=0.999)
    ema_ref = [p.detach().clone() for p in ma.parameters()]
    for it in range(3):
        # the gradients come from the fused model's real training step and are handed to both optimizers (Adam's update
        # m/sqrt(v) amplifies rounding noise of near-zero gradients to +-lr, so each needs the SAME gradients)
        torch.manual_seed(100 + it)
        ob.zero_grad(set_to_none=True)
        loss = mb.loss(x, mask=mask, cond=None)
        loss.backward()
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            pa.grad = pb.grad.detach().clone()
        norm = float(torch.nn.utils.clip_grad_norm_(ma.parameters(), 0.5))
        oa.step()
        ob.step()
        for e, p_ in zip(ema_ref, ma.parameters()):
            diff = e - p_.detach()
            diff.mul_(1.0 - 0.999)
            e.sub_(diff)
        assert abs(float(ob.last_grad_norm) - norm) <= 1e-5 * norm
        for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
            assert torch.allclose(pa, pb, rtol=1e-5, atol=2e-7), (it, k, float((pa - pb).abs().max()))
    for e, eb in zip(ema_ref, ob.ema_parameters()):
        assert torch.allclose(e, eb, rtol=1e-5, atol=1e-7)
    # the parameters live in one flat buffer laid out like the library's flat gradient
    ptrs = sorted(p.data_ptr() for p in mb.parameters())
    assert ptrs[-1] - ptrs[0] < 4 * sum(p.numel() for p in mb.parameters())


def _fixed_batch():
    g = Golden("c1_jetnet30")
    gen = torch.Generator().manual_seed(3)
    B, N = 64, 30
    n_real = torch.randint(5, N + 1, (B,), generator=gen)
    mask = (torch.arange(N)[None, :] < n_real[:, None]).float().unsqueeze(-1)
    x = torch.randn(B, N, 3, generator=gen) * 5.0 * mask
    return g, x.to(DEV), mask.to(DEV)


def test_training_with_the_fused_optimizer_learns():
    """End to end: fused loss + FusedClipAdamW for 40 steps on one batch with the SAME draws every step -- the loss must go
    down, i.e. the optimizer's raw-pointer parameter update is seen by the packed weight copies of the next forward."""
    from particle_fm_b200.optim import FusedClipAdamW
    g, x, mask = _fixed_batch()
    m = build_module(g.ctor, g.sd, device=DEV)
    opt = FusedClipAdamW(m.parameters(), lr=2e-3, weight_decay=0.0, max_grad_norm=0.5)
    losses = []
    for it in range(40):
        torch.manual_seed(11)
        opt.zero_grad(set_to_none=True)
        loss = m.loss(x, mask=mask, cond=None)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.9 * losses[0], losses[::8]
    # clipped AdamW on one batch is not strictly monotone (atomics reorder the fp32 sums from run to run): no step may jump
    assert all(b <= a * 1.15 for a, b in zip(losses, losses[1:])), losses[::4]
    assert sum(losses[-5:]) / 5 < 0.9 * losses[0]


def test_graphed_training_step_learns_and_matches_eager_shapes():
    """launch.GraphedTrainStep: the whole step replayed from a CUDA graph (per-jet times drawn on the CPU generator outside
    the graph).  Over 60 replays on one batch the loss decreases; parameters keep living in the optimizer's flat buffer."""
    from particle_fm_b200.launch import GraphedTrainStep
    from particle_fm_b200.optim import FusedClipAdamW
    g, x, mask = _fixed_batch()
    m = build_module(g.ctor, g.sd, device=DEV)
    opt = FusedClipAdamW(m.parameters(), lr=2e-3, weight_decay=0.0, max_grad_norm=0.5, device_step_count=True)
    step = GraphedTrainStep(m, opt, x, mask)
    p0 = [p.detach().clone() for p in m.parameters()]
    losses = [float(step(x, mask)) for _ in range(300)]       # fresh noise every replay: compare 30-step averages
    first, last = sum(losses[:30]) / 30, sum(losses[-30:]) / 30
    assert last < 0.85 * first, (first, last)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(p0, m.parameters()))
    # the graph and the eager path leave the same kind of state behind: an eager evaluation still works and sees the new weights
    torch.manual_seed(1)
    with torch.no_grad():
        l_eager = float(m.loss(x, mask=mask, cond=None))
    assert abs(l_eager - last) < 0.35 * last, (l_eager, last)


def test_graphed_training_step_after_an_eager_backward_on_the_default_stream():
    """An eager backward on the default stream whose ``loss`` is still alive keeps the parameters' AccumulateGrad nodes bound to
    the legacy stream; the captured step must not run them (the library writes p.grad itself, training.direct_param_grads)."""
    from particle_fm_b200.launch import GraphedTrainStep
    from particle_fm_b200.optim import FusedClipAdamW
    g, x, mask = _fixed_batch()
    m = build_module(g.ctor, g.sd, device=DEV)
    kept = m.loss(x[:7], mask=mask[:7], cond=None)
    kept.backward()                                        # grad accumulators created on the default stream, graph kept alive
    opt = FusedClipAdamW(m.parameters(), lr=2e-3, weight_decay=0.0, max_grad_norm=0.5, device_step_count=True)
    step = GraphedTrainStep(m, opt, x, mask)
    first = float(step(x, mask))
    for _ in range(60):
        last = float(step(x, mask))
    assert first == first and last == last and float(kept) == float(kept)
    assert last < first

