"""Drop-in for particle_fm/models/flow_matching_module.py on B200.

``ode_wrapper`` (:34-71), ``CNF`` (:74-347) and ``SetFlowMatchingLitModule`` (:350-677) keep the
reference's constructor arguments, attribute names (``flows``, ``loss``, ``conditioned``,
``hparams``), state_dict keys and method signatures, so Hydra configs
(``_target_: ...flow_matching_module.SetFlowMatchingLitModule``), checkpoints, the EMA callback and
``generate_data`` keep working.  Underneath, ``sample()``/``forward(reverse=True)`` run the whole
Euler / midpoint integration as one fused CUDA launch and ``training_step`` runs the fused
flow-matching loss forward + backward of libpfm_b200.so.  Options the CUDA path does not implement
raise ``NotImplementedError`` instead of silently diverging -- there is no PyTorch fallback.
"""
from __future__ import annotations

import inspect
import types
from typing import Any, Mapping, Optional

import torch
import torch.nn as nn
from torch import Tensor

from .components.epic import EPiC_encoder
from .components.losses import (ConditionalFlowMatchingLoss, ConditionalFlowMatchingOTLoss, DiffusionLoss, DroidLoss,
                                FlowMatchingLoss)
from .components.time_emb import CosineEncoding, sincos_encoding

try:  # Lightning is optional: it is not installed in the build image
    import pytorch_lightning as pl
    _LightningBase = pl.LightningModule
except Exception:  # pragma: no cover - depends on the environment
    try:
        import lightning.pytorch as pl
        _LightningBase = pl.LightningModule
    except Exception:
        pl = None

        class _LightningBase(nn.Module):
            """Minimal stand-in for pl.LightningModule: hparams, device, log()."""

            def save_hyperparameters(self, *args, **kwargs):
                frame = inspect.currentframe().f_back
                names, _, _, values = inspect.getargvalues(frame)
                self.hparams = types.SimpleNamespace(**{n: values[n] for n in names if n != "self"})

            @property
            def device(self) -> torch.device:
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def log(self, *args, **kwargs):
                pass

            current_epoch = 0
            trainer = None

FIXED_STEP_SOLVERS = ("euler", "midpoint")


def fixed_step_grid(ode_steps: int, solver: str):
    """fp32 evaluation times and step sizes of ``NeuralODE(..., solver).trajectory(z, linspace(1, 0, ode_steps))``
    (flow_matching_module.py:278-287).  torchdyn integrates the reversed grid -linspace(1,0,n) with
    f_(t,x) = -f(-t,x), keeps t by accumulation (t += dt; dt = t_span[k+1] - t) and evaluates the midpoint
    at t + 0.5*dt; the cosine code is chaotic in t, so the recurrence is reproduced operation by operation
    in fp32 torch scalars.  Returns (t_eval [n_evals], dt [ode_steps-1]) on the CPU."""
    if ode_steps < 2:
        raise ValueError("ode_steps must be >= 2 (it counts grid points, not steps)")
    t_span = -torch.linspace(1.0, 0.0, ode_steps)
    t = t_span[0]
    dt = t_span[1] - t_span[0]
    t_eval, dts = [], []
    n = ode_steps - 1
    for k in range(1, n + 1):
        t_eval.append(-t)
        if solver == "midpoint":
            t_eval.append(-(t + 0.5 * dt))
        dts.append(dt)
        t = t + dt
        if k < n:
            dt = t_span[k + 1] - t
    return torch.stack(t_eval), torch.stack(dts)


class ode_wrapper(nn.Module):
    """Kept for API compatibility (flow_matching_module.py:34-71): binds cond / mask to the vector field."""

    def __init__(self, model: nn.Module, mask: Tensor = None, cond: Tensor = None, loss_type: str = "FM-OT",
                 diff_config: Mapping = {"max_sr": 0.999, "min_sr": 0.02}):
        super().__init__()
        self.model, self.mask, self.cond, self.loss_type = model, mask, cond, loss_type
        if loss_type == "diffusion":
            from .components.diffusion import VPDiffusionSchedule
            self.diff_sched = VPDiffusionSchedule(**diff_config)

    def forward(self, t, x, *args, **kwargs):
        if self.loss_type == "diffusion":      # probability-flow ODE drift (:62-69); sampling uses the fused program instead
            shape = [-1] + [1] * (x.dim() - 1)
            _, noise_rates = self.diff_sched(t.view(shape))
            betas = self.diff_sched.get_betas(t.view(shape))
            return -0.5 * betas * (x - self.model(t, x, mask=self.mask, cond=self.cond) / noise_rates)
        return self.model(t, x, mask=self.mask, cond=self.cond)


class CNF(nn.Module):
    """Continuous normalizing flow around the EPiC vector field (flow_matching_module.py:74-347)."""

    def __init__(self, model: str = "epic", features: int = 3, num_particles: int = 150, frequencies: int = 6,
                 hidden_dim: int = 128, layers: int = 8, global_cond_dim: int = 0, local_cond_dim: int = 0,
                 dropout: float = 0.0, latent: int = 16, activation: str = "leaky_relu",
                 wrapper_func: str = "weight_norm", t_local_cat: bool = False, t_global_cat: bool = False,
                 add_time_to_input: bool = True, t_emb: str = "sincos", loss_type: str = "FM-OT",
                 diff_config: Mapping[str, Any] = {"max_sr": 0.999, "min_sr": 0.02}, sum_scale: float = 1e-2,
                 net_config: Mapping[str, Any] = {}):
        super().__init__()
        self.latent = latent
        self.features = features
        self.add_time_to_input = add_time_to_input
        input_dim = features + 2 * frequencies if self.add_time_to_input else features
        if model == "epic":
            self.net = EPiC_encoder(input_dim=input_dim, feats=features, latent=latent, equiv_layers=layers,
                                    hid_d=hidden_dim, activation=activation, wrapper_func=wrapper_func,
                                    frequencies=frequencies, num_points=num_particles, t_local_cat=t_local_cat,
                                    t_global_cat=t_global_cat, global_cond_dim=global_cond_dim,
                                    local_cond_dim=local_cond_dim, dropout=dropout, sum_scale=sum_scale)
        elif model in ("droid_fulltransformer", "droid_fullcrossattention"):
            from .components.droid_transformer import FullCrossAttentionEncoder, FullTransformerEncoder
            cls = FullTransformerEncoder if model == "droid_fulltransformer" else FullCrossAttentionEncoder
            self.net = cls(inpt_dim=input_dim, outp_dim=features, ctxt_dim=global_cond_dim + 2 * frequencies,
                           **dict(net_config))
            self.net.t_dim = 2 * frequencies
        else:
            raise NotImplementedError(f"Model {model} not implemented.")
        self.register_buffer("frequencies", 2 ** torch.arange(frequencies) * torch.pi)
        self.n_frequencies = frequencies
        self.activation = activation
        self.t_emb = t_emb
        self.loss_type = loss_type
        self.diff_config = diff_config
        if self.t_emb == "cosine":
            self.embed = CosineEncoding(outp_dim=2 * frequencies, min_value=0.0, max_value=1.0,
                                        frequency_scaling="exponential")
        elif self.t_emb != "sincos":
            raise NotImplementedError(f"t_emb={t_emb!r}: the CUDA path implements 'cosine' and 'sincos'")

    # -- time codes ----------------------------------------------------------------------------
    def time_code(self, t: Tensor) -> Tensor:
        """(..., 2*frequencies) code of times t (any shape); same ops as CNF.time_embedding :206-233."""
        if self.t_emb == "sincos":
            return sincos_encoding(t, self.frequencies.to(t.device))
        if t.dim() == 0:
            t = t.unsqueeze(0)
        return self.embed(t)

    def time_embedding(self, t: Tensor, x: Tensor, t_emb: str = "sincos") -> Tensor:
        return self.time_code(t).expand(*x.shape[:-1], -1)

    def forward(self, t: Tensor, x: Tensor, cond: Tensor = None, mask: Tensor = None) -> Tensor:
        """v(t, x): t is 0-dim (sampling) or (B, N) (training); x (B, N, features)."""
        code = self.time_embedding(t, x, self.t_emb)
        if self.add_time_to_input:
            x = torch.cat((code, x), dim=-1)
        return self.net(code, x, cond, mask)

    # -- sampling ------------------------------------------------------------------------------
    def encode(self, x: Tensor, mask: Tensor = None, ode_solver: str = "dopri5_zuko", ode_steps: int = 100) -> Tensor:
        raise NotImplementedError("CNF.encode (data -> noise, rk4) is not part of the B200 hot path")

    @torch.no_grad()
    def decode(self, z: Tensor, cond: Tensor, mask: Tensor = None, ode_solver: str = "dopri5_zuko",
               ode_steps: int = 100) -> Tensor:
        """Integrate from t=1 (noise) to t=0 (data) -- one fused CUDA launch for all steps."""
        if self.loss_type == "diffusion":
            return self._decode_diffusion(z, cond, mask, ode_solver, ode_steps)
        if ode_solver in ("em", "ddim"):
            raise SyntaxError(f"Solver {ode_solver} is only implemented for diffusion loss")      # :328-329
        if ode_solver not in FIXED_STEP_SOLVERS:
            raise NotImplementedError(f"ode_solver={ode_solver!r}: the B200 path implements the fixed-step "
                                      f"{FIXED_STEP_SOLVERS} solvers (the reference's generation configs use midpoint)")
        key = (ode_steps, ode_solver)
        cache = self.__dict__.setdefault("_grid_cache", {})
        if key not in cache:
            t_eval, dt = fixed_step_grid(ode_steps, ode_solver)
            cache[key] = (self.time_code(t_eval), dt)  # [n_evals, T], evaluated on the CPU like the oracle
        codes, dt = cache[key]
        if z.device.type == "cuda":               # device copies cached too: a pageable host -> device copy per call would
            dkey = (key, z.device)                #  make the host wait for the GPU work already queued on the stream
            if dkey not in cache:
                cache[dkey] = (codes.to(z.device), dt.to(z.device))
            codes, dt = cache[dkey]
        # one fold + repack launch is negligible next to a whole integration: re-sync unconditionally, so that in-place
        # parameter updates the (_version, data_ptr) key cannot see (p.data.mul_(), manual EMA) never sample stale weights
        eng = self.net.engine(force_sync=True)
        takes = self.net.t_local_cat or self.net.t_global_cat
        return eng.sample(z, mask, cond, codes if takes else None, codes if self.add_time_to_input else None, dt,
                          ode_solver)


def _diffusion_program(cnf, ode_solver: str, ode_steps: int):
    """Per-evaluation (times, coefficient rows, dt) of the diffusion samplers, computed on the CPU in fp32 with the
    reference's own recurrences so that the time codes and schedule values round the same way:
      ddim  solver.py:60-92   diff_times = 1, then ``diff_times - step_size`` per step; rates at t and at t - step
      em    solver.py:112-133 t = 1, ``t -= delta_t``; betas(t), noise_rate(t), delta_t, sqrt(betas * delta_t)
      euler / midpoint        the torchdyn grid of ``fixed_step_grid`` with the drift of ode_wrapper.forward :62-69"""
    from .components.diffusion import VPDiffusionSchedule
    sched = VPDiffusionSchedule(**cnf.diff_config)
    if ode_solver == "ddim":
        step = 1 / ode_steps
        t = torch.ones(1)
        nsr, nnr = sched(t.view(-1, 1, 1))
        times, rows = [], []
        for _ in range(ode_steps):
            sr, nr = nsr, nnr
            times.append(t[0].clone())
            t = t - step
            nsr, nnr = sched(t.view(-1, 1, 1))
            rows.append(torch.stack([sr.reshape(()), nr.reshape(()), nsr.reshape(()), nnr.reshape(())]))
        return torch.stack(times), torch.stack(rows), None
    if ode_solver == "em":
        delta_t = 1 / ode_steps
        t = torch.ones(1)
        times, rows = [], []
        for _ in range(ode_steps):
            times.append(t[0].clone())
            _, nr = sched(t.view(-1, 1, 1))
            betas = sched.get_betas(t.view(-1, 1, 1))
            rows.append(torch.stack([betas.reshape(()), nr.reshape(()), torch.tensor(delta_t, dtype=torch.float32),
                                     (betas * delta_t).sqrt().reshape(())]))
            t = t - delta_t                      # the reference's in-place ``t -= delta_t``: same fp32 arithmetic
        return torch.stack(times), torch.stack(rows), None
    t_eval, dt = fixed_step_grid(ode_steps, ode_solver)
    tv = t_eval.view(-1, 1, 1)
    _, nr = sched(tv)
    betas = sched.get_betas(tv)
    z = torch.zeros_like(t_eval)
    return t_eval, torch.stack([betas.reshape(-1), nr.reshape(-1), z, z], dim=1), dt


def _cnf_decode_diffusion(self, z, cond, mask, ode_solver, ode_steps):
    """CNF.decode for loss_type == 'diffusion' (flow_matching_module.py:245-329): DDIM / Euler-Maruyama samplers and the
    probability-flow ODE (euler / midpoint) as step programs of the single-launch integrator."""
    if ode_solver not in ("em", "ddim") + FIXED_STEP_SOLVERS:
        raise NotImplementedError(f"ode_solver={ode_solver!r} with the diffusion loss: the B200 path implements "
                                  f"'em', 'ddim' and the fixed-step {FIXED_STEP_SOLVERS} probability-flow ODE")
    from .components.epic import EPiC_encoder
    if not isinstance(self.net, EPiC_encoder):
        raise NotImplementedError("the diffusion samplers of the B200 path run on the EPiC vector field")
    key = ("diffusion", ode_steps, ode_solver)
    cache = self.__dict__.setdefault("_grid_cache", {})
    if key not in cache:
        t_eval, coef, dt = _diffusion_program(self, ode_solver, ode_steps)
        cache[key] = (self.time_code(t_eval), coef, dt)
    codes, coef, dt = cache[key]
    eng = self.net.engine(force_sync=True)
    if eng.precision != "fp32":
        eng.set_precision("fp32")             # the step programs live in the fp32 kernel
        self.net.precision = "fp32"
    takes = self.net.t_local_cat or self.net.t_global_cat
    noise = None
    if ode_solver == "em":
        # solver.py:131 draws torch.randn_like(x_t) on the device at every step: same generator, same call sequence
        noise = torch.stack([torch.randn_like(z) for _ in range(ode_steps)])
    kind = {"em": "em", "ddim": "ddim"}.get(ode_solver, "pf_ode")
    return eng.sample_diffusion(z, mask, cond, codes if takes else None, codes if self.add_time_to_input else None, coef,
                                kind, solver=ode_solver if kind == "pf_ode" else "euler", dt=dt, noise=noise)


CNF._decode_diffusion = _cnf_decode_diffusion


class SetFlowMatchingLitModule(_LightningBase):
    """LightningModule for set flow matching (flow_matching_module.py:350-677), B200 hot path underneath."""

    def __init__(self, optimizer: torch.optim.Optimizer, scheduler: torch.optim.lr_scheduler = None,
                 model: str = "epic", features: int = 3, hidden_dim: int = 128, num_particles: int = 150,
                 frequencies: int = 6, layers: int = 8, n_transforms: int = 1, activation: str = "leaky_relu",
                 wrapper_func: str = "weight_norm", use_normaliser: bool = False, normaliser_config: Mapping = {},
                 net_config: Mapping = {},
                 # epic
                 latent: int = 16, t_local_cat: bool = False, t_global_cat: bool = False,
                 add_time_to_input: bool = True, global_cond_dim: int = 0, local_cond_dim: int = 0,
                 dropout: float = 0.0, sum_scale: float = 1e-2,
                 # loss
                 loss_type: str = "FM-OT", sigma: float = 1e-4, t_emb: str = "sincos",
                 diff_config: Mapping = {"max_sr": 1, "min_sr": 1e-8}, criterion: str = "mse"):
        super().__init__()
        self.save_hyperparameters(logger=False)
        if use_normaliser:
            raise NotImplementedError("use_normaliser=True is not supported by the B200 path (False in every model YAML)")
        if n_transforms != 1:
            raise NotImplementedError("n_transforms != 1 is not supported by the B200 path (1 in every model YAML)")
        flows = nn.ModuleList()
        for _ in range(n_transforms):
            flows.append(CNF(model=model, net_config=net_config, features=features, hidden_dim=hidden_dim,
                             num_particles=num_particles, frequencies=frequencies, layers=layers,
                             global_cond_dim=global_cond_dim, local_cond_dim=local_cond_dim, latent=latent,
                             dropout=dropout, activation=activation, wrapper_func=wrapper_func,
                             t_global_cat=t_global_cat, t_local_cat=t_local_cat, add_time_to_input=add_time_to_input,
                             t_emb=t_emb, loss_type=loss_type, diff_config=diff_config, sum_scale=sum_scale))
        self.flows = flows
        self.conditioned = global_cond_dim > 0
        if loss_type == "FM-OT":
            self.loss = FlowMatchingLoss(flows=self.flows, sigma=sigma, criterion=criterion)
        elif loss_type == "CFM":
            self.loss = ConditionalFlowMatchingLoss(flows=self.flows, sigma=sigma, criterion=criterion)
        elif loss_type == "CFM-OT":
            self.loss = ConditionalFlowMatchingOTLoss(flows=self.flows, sigma=sigma, criterion=criterion)
        elif loss_type == "diffusion":
            self.loss = DiffusionLoss(flows=self.flows, sigma=sigma, diff_config=diff_config, criterion=criterion)
        elif loss_type == "droid":
            self.loss = DroidLoss(flows=self.flows, sigma=sigma, criterion=criterion)
        else:
            raise NotImplementedError(f"Loss type {loss_type} not implemented.")

    # -- precision knob of the CUDA path (not in the reference) ---------------------------------
    def set_precision(self, precision: str):
        """'fp32' (CUDA cores, strict parity) or 'bf16' (tcgen05 tensor cores, 2e-2 per-step parity)."""
        for f in self.flows:
            f.net.precision = precision
        return self

    def forward(self, x: Tensor, cond: Tensor = None, mask: Tensor = None, reverse: bool = False,
                ode_solver: str = "dopri5_zuko", ode_steps: int = 100):
        if reverse:
            for f in reversed(self.flows):
                x = f.decode(x, cond, mask, ode_solver=ode_solver, ode_steps=ode_steps)
        else:
            for f in self.flows:
                x = f.encode(x, mask, ode_solver=ode_solver, ode_steps=ode_steps)
        return x

    def _variable_jet_sizes(self) -> bool:
        dm = getattr(getattr(self, "trainer", None), "datamodule", None)
        if dm is None:
            return True
        return bool(dm.hparams.variable_jet_sizes)

    def training_step(self, batch, batch_idx):
        x, mask, cond = batch
        if not self._variable_jet_sizes():       # flow_matching_module.py:519-520
            mask = None
        loss = self.loss(x, mask=mask, cond=cond)
        self.log("train/loss", loss, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        # per-jet-type losses every 20 epochs (flow_matching_module.py:526-551, JetClass-cond datamodule option); the extra
        # loss calls consume the RNG streams exactly like the reference's (rand(B) on the CPU, randn_like on the device)
        dm = getattr(getattr(self, "trainer", None), "datamodule", None)
        if dm is not None and self.current_epoch % 20 == 0 and hasattr(dm.hparams, "loss_per_jettype"):
            if dm.hparams.loss_per_jettype:
                names = list(dm.names_conditioning)
                for jet_type in dm.hparams.used_jet_types:
                    sel = cond[:, names.index(f"jet_type_label_{jet_type}")] == 1
                    x_t, m_t, c_t = x[sel][:10_000], mask[sel][:10_000], cond[sel][:10_000]
                    if x_t.shape[0] == 0:              # the reference divides 0 by 0 here
                        loss_t = torch.full((), float("nan"), device=x.device)
                    else:
                        with torch.no_grad():          # logging only: the reference never back-propagates these
                            loss_t = self.loss(x_t, m_t, cond=c_t)
                    self.log(f"train/loss_{jet_type}", loss_t, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        return {"loss": loss}

    def on_validation_epoch_start(self) -> None:
        torch.manual_seed(9999)                  # :555-557 same seed for every validation epoch

    def on_validation_epoch_end(self) -> None:
        torch.manual_seed(torch.seed())

    def validation_step(self, batch: Any, batch_idx: int):
        x, mask, cond = batch
        if not self._variable_jet_sizes():
            mask = None
        with torch.no_grad():
            loss = self.loss(x, mask, cond=cond)
        self.log("val/loss", loss, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        return {"loss": loss}

    def test_step(self, batch: Any, batch_idx: int):
        pass

    def configure_optimizers(self):
        optimizer = self.hparams.optimizer(params=self.parameters())
        if self.hparams.scheduler is not None:
            scheduler = self.hparams.scheduler(optimizer=optimizer)
            return {"optimizer": optimizer,
                    "lr_scheduler": {"scheduler": scheduler, "monitor": "val/loss", "interval": "epoch", "frequency": 1}}
        return {"optimizer": optimizer}

    def __getstate__(self):             # transfer helpers (pinned buffer, event, copy stream) are per-process: never pickled / deep-copied
        st = self.__dict__.copy()
        for k in ("_noise_pin", "_noise_evt", "_copy_stream"):
            st.pop(k, None)
        return st

    @torch.no_grad()
    def sample(self, n_samples: int, cond: Tensor = None, mask: Tensor = None, ode_solver: str = "midpoint",
               ode_steps: int = 100, num_points: int = None):
        """Generate samples (flow_matching_module.py:637-677): noise from the CPU default generator
        (same stream as the reference), masked, integrated 1 -> 0 on the GPU in one launch."""
        shape = (n_samples, num_points if num_points else self.hparams.num_particles, self.hparams.features)
        if cond is not None:
            cond = cond.to(self.device)
        if mask is not None:
            mask = mask[:n_samples]
            mask = mask.to(self.device)
        if self.device.type != "cuda":
            z = torch.randn(shape).to(self.device)
            if mask is not None:
                z = z * mask
            return self.forward(z, cond=cond, mask=mask, reverse=True, ode_solver=ode_solver, ode_steps=ode_steps)
        # The noise comes from the CPU default generator like the reference's torch.randn(shape), drawn straight into a
        # cached pinned buffer and copied on a side stream: nothing here waits for GPU work already queued, so in a loop of
        # calls (generate_data batches) the draw for call k+1 runs while the GPU integrates call k.
        numel = n_samples * shape[1] * shape[2]
        buf = self.__dict__.get("_noise_pin")
        if buf is None or buf.numel() < numel:
            buf = torch.empty(numel, pin_memory=True)
            self.__dict__["_noise_pin"] = buf
        evt = self.__dict__.get("_noise_evt")
        if evt is not None:
            evt.synchronize()                  # the previous call's copy out of this buffer has completed
        zc = torch.randn(shape, out=buf[:numel].view(shape))
        main = torch.cuda.current_stream(self.device)
        side = self.__dict__.get("_copy_stream")
        if side is None or side.device != self.device:
            side = torch.cuda.Stream(device=self.device)
            self.__dict__["_copy_stream"] = side
        with torch.cuda.stream(side):
            z = zc.to(self.device, non_blocking=True)
            evt = torch.cuda.Event()
            evt.record(side)
        self.__dict__["_noise_evt"] = evt
        main.wait_event(evt)
        z.record_stream(main)
        if mask is not None:
            z = z * mask
        return self.forward(z, cond=cond, mask=mask, reverse=True, ode_solver=ode_solver, ode_steps=ode_steps)
