"""Drop-in for particle_fm/models/flow_matching_no_sets.py on B200 -- the jet-feature flow (LHCO step 1).

``CNF`` (:41-112) and ``FLowMatchingNoSetsLitModule`` (:115-238) keep the reference's constructor arguments,
attributes (``flows``, ``loss``) and state_dict keys (``flows.0.net.mlp1.0.weight`` ..., ``flows.0.freqs``).
``sample()`` / ``forward(reverse=True)`` integrate all midpoint steps in ONE launch of ``pfm_mlp_sample`` (state of a
row tile resident in shared memory); the result stays on the device, so the LHCO chain
(``particle_fm_b200.launch.lhco_chain``) can hand it to the particle model without a host round trip."""
from __future__ import annotations

import torch
from torch import nn

from .components.mlp import small_cond_MLP_model
from .flow_matching_module import FIXED_STEP_SOLVERS, _LightningBase, fixed_step_grid


class ode_wrapper(nn.Module):
    """flow_matching_no_sets.py:17-38 (API compatibility)."""

    def __init__(self, model: nn.Module, mask: torch.Tensor = None, cond: torch.Tensor = None):
        super().__init__()
        self.model, self.mask, self.cond = model, mask, cond

    def forward(self, t, x, *args, **kwargs):
        return self.model(t, x, mask=self.mask, cond=self.cond)


class CNF(nn.Module):
    def __init__(self, features: int, freqs: int = 3, activation: str = "Tanh"):
        super().__init__()
        self.net = small_cond_MLP_model(features, features, dim_t=2 * freqs, dim_cond=1, activation=activation)
        self.register_buffer("freqs", torch.arange(1, freqs + 1) * torch.pi)

    def time_code(self, t: torch.Tensor) -> torch.Tensor:
        t = self.freqs.to(t.device) * t[..., None]            # :62-63
        return torch.cat((t.cos(), t.sin()), dim=-1)

    def forward(self, t: torch.Tensor, x: torch.Tensor, mask: torch.Tensor = None, cond: torch.Tensor = None) -> torch.Tensor:
        code = self.time_code(t)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.net.parameters()) or x.requires_grad:
            code = code.expand(*x.shape[:-1], -1)             # :64
        elif code.dim() == 1:
            code = code.unsqueeze(0)                          # one code for the whole batch: the kernel broadcasts it
        return self.net(code, x, cond=cond)

    def encode(self, x: torch.Tensor, *a, **k) -> torch.Tensor:
        raise NotImplementedError("CNF.encode (data -> noise) is not part of the B200 hot path")

    @torch.no_grad()
    def decode(self, z: torch.Tensor, cond: torch.Tensor, mask: torch.Tensor = None, ode_solver: str = "midpoint",
               ode_steps: int = 100) -> torch.Tensor:
        if ode_solver not in FIXED_STEP_SOLVERS:
            raise NotImplementedError(f"Solver {ode_solver} not implemented")      # the reference only has midpoint (:91-92)
        key = (ode_steps, ode_solver)
        cache = self.__dict__.setdefault("_grid_cache", {})
        if key not in cache:
            t_eval, dt = fixed_step_grid(ode_steps, ode_solver)
            cache[key] = (self.time_code(t_eval), dt)
        codes, dt = cache[key]
        return self.net.engine(z.device, force_sync=True).sample(z, cond, codes, dt, ode_solver)

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("CNF.log_prob (adaptive augmented ODE) is not part of the B200 hot path")


class FLowMatchingNoSetsLitModule(_LightningBase):
    def __init__(self, optimizer: torch.optim.Optimizer, scheduler: torch.optim.lr_scheduler = None, features: int = 10,
                 n_transforms: int = 1, sigma: float = 1e-4, activation: str = "ELU", freqs: int = 3):
        super().__init__()
        self.save_hyperparameters(logger=False)
        if n_transforms != 1:
            raise NotImplementedError("n_transforms != 1 is not supported by the B200 path")
        self.flows = nn.ModuleList([CNF(features, freqs=freqs, activation=activation) for _ in range(n_transforms)])
        self.loss = _RowFlowMatchingLoss(flows=self.flows, sigma=sigma)

    def forward(self, x: torch.Tensor, cond: torch.Tensor = None, mask: torch.Tensor = None, reverse: bool = False,
                ode_solver: str = "midpoint", ode_steps: int = 100):
        if reverse:
            for f in reversed(self.flows):
                x = f.decode(x, cond, mask, ode_solver=ode_solver, ode_steps=ode_steps)
        else:
            for f in self.flows:
                x = f.encode(x, mask, ode_solver=ode_solver, ode_steps=ode_steps)
        return x

    def training_step(self, batch, batch_idx):
        x, mask, cond = batch
        loss = self.loss(x, cond=cond)
        self.log("train/loss", loss, on_step=False, on_epoch=True, prog_bar=True)
        return {"loss": loss}

    def on_validation_epoch_start(self) -> None:
        torch.manual_seed(9999)

    def on_validation_epoch_end(self) -> None:
        torch.manual_seed(torch.seed())

    def validation_step(self, batch, batch_idx: int):
        x, mask, cond = batch
        loss = self.loss(x, cond=cond)
        self.log("val/loss", loss, on_step=False, on_epoch=True, prog_bar=True)
        return {"loss": loss}

    def test_step(self, batch, batch_idx: int):
        pass

    def configure_optimizers(self):
        optimizer = self.hparams.optimizer(params=self.parameters())
        if self.hparams.scheduler is not None:
            scheduler = self.hparams.scheduler(optimizer=optimizer)
            return {"optimizer": optimizer,
                    "lr_scheduler": {"scheduler": scheduler, "monitor": "val/loss", "interval": "epoch", "frequency": 1}}
        return {"optimizer": optimizer}

    @torch.no_grad()
    def sample(self, n_samples: int, mask: torch.Tensor = None, cond: torch.Tensor = None, ode_solver: str = "midpoint",
               ode_steps: int = 100):
        """flow_matching_no_sets.py:212-238: z ~ N(0,1) on the CPU generator, moved, integrated 1 -> 0 on the GPU."""
        z = torch.randn(n_samples, self.hparams.features).to(self.device)
        if cond is not None:
            cond = cond.to(self.device)
        return self.forward(z, cond=cond, mask=mask, reverse=True, ode_solver=ode_solver, ode_steps=ode_steps)


class _RowFlowMatchingLoss(nn.Module):
    """FlowMatchingLoss.forward on 2-D rows (losses.py:38-77, the ``else`` branch of :44-50: t per row on x's device)."""

    def __init__(self, flows: nn.ModuleList, sigma: float = 1e-4):
        super().__init__()
        self.flows, self.sigma = flows, sigma

    def forward(self, x: torch.Tensor, mask: torch.Tensor = None, cond: torch.Tensor = None, draws=None) -> torch.Tensor:
        if x.dim() != 2:
            raise NotImplementedError("the jet-feature flow trains on rows (B, features)")
        if mask is None:
            mask = torch.ones_like(x[..., 0]).unsqueeze(-1)
        if draws is None:
            t = torch.rand_like(x[..., 0]).unsqueeze(-1)
            z = torch.randn_like(x)
        else:
            t, z = draws
        y = (1 - t) * x + (self.sigma + (1 - self.sigma) * t) * z
        u_t = ((1 - self.sigma) * z - x) * mask
        temp = y.clone()
        for v in self.flows:
            temp = v(t.squeeze(-1), temp, mask=mask, cond=cond)
        return (temp - u_t).square().sum() / mask.sum()
