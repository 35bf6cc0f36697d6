"""Oracle (test infrastructure): CPU restatements of the scope table's "next" rows.

  post_process            particle_fm/utils/data_generation.py:105-123 (+ data/components/utils.py:183-200)
  cfm_ot_loss             ConditionalFlowMatchingOTLoss.forward, models/components/losses.py:147-204
  diffusion_loss          DiffusionLoss.forward, losses.py:230-285
  ddim_sample / em_sample components/solver.py:23-95, :98-143
  pf_ode_sample           CNF.decode euler/midpoint with loss_type == "diffusion": ode_wrapper.forward
                          flow_matching_module.py:62-69 under the torchdyn loop of oracle/ode_oracle.py
  mlp_flow_*              small_cond_MLP_model (components/mlp.py:24-68) behind CNF of
                          models/flow_matching_no_sets.py:41-93 and FlowMatchingLoss on rows (losses.py:44-50)
All random draws are injectable so that the CUDA path can be fed the same numbers.  Pinned against the UNMODIFIED
reference by oracle/make_golden_next.py (same seeds -> identical results), except:
  * PARITY UNPINNED: the optimal-transport plan is third-party POT (``ot.emd``, unpinned requirements.txt:29), absent
    here.  With uniform marginals of equal size the plan is an assignment; ``scipy.optimize.linear_sum_assignment``
    solves the same problem exactly.  The reference's forward is exercised with that stand-in for ``ot.emd``.
  * the ODE grid is the restated torchdyn loop (oracle/ode_oracle.py, parity unpinned as stated there).
"""
from __future__ import annotations

import math
from typing import Callable, List, Mapping, Optional, Sequence

import numpy as np
import torch

from . import ode_oracle

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# generate_data post-processing
# ----------------------------------------------------------------------------------------------
def inverse_normalize_tensor(tensor, mean, std, sigma=5):
    for i in range(len(mean)):                                   # utils.py:198-199
        tensor[..., i] = (tensor[..., i] * (std[i] / sigma)) + mean[i]
    return tensor


def post_process(batch: Tensor, mask_batch, normalized_data, normalize_sigma, means, stds, log_pt, pt_standardization,
                 variable_set_sizes) -> Tensor:
    """What generate_data does to one sampled batch after ``.cpu()`` (data_generation.py:105-122)."""
    batch = batch.clone()
    if normalized_data:
        if pt_standardization:
            batch[..., :2] = inverse_normalize_tensor(batch[..., :2], means[:2], stds[:2], sigma=10)
            batch[..., 2] = inverse_normalize_tensor(batch[..., 2], [means[2]], [stds[2]], sigma=5)
        else:
            batch = inverse_normalize_tensor(batch, means, stds, sigma=normalize_sigma)
        if log_pt:
            batch[..., 2] = 1.0 - np.exp(batch[..., 2])
    if variable_set_sizes:
        batch = batch * mask_batch
    return batch


# ----------------------------------------------------------------------------------------------
# CFM-OT
# ----------------------------------------------------------------------------------------------
def emd_uniform(M: np.ndarray) -> np.ndarray:
    """Stand-in for ``ot.emd(unif(n), unif(n), M)``: the optimal plan of a square problem with uniform marginals is
    a permutation matrix / n (a vertex of the Birkhoff polytope)."""
    from scipy.optimize import linear_sum_assignment
    n = M.shape[0]
    r, c = linear_sum_assignment(M.astype(np.float64))
    pi = np.zeros((n, n))
    pi[r, c] = 1.0 / n
    return pi


def choice_from_uniform(p: np.ndarray, u: np.ndarray) -> np.ndarray:
    """numpy's legacy ``RandomState.choice(len(p), p=p, size=len(u))`` given its uniform draws u
    (numpy/random/mtrand.pyx: cdf = p.cumsum(); cdf /= cdf[-1]; cdf.searchsorted(u, side='right'))."""
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(u, side="right")


def cfm_ot_loss(vf: Callable, x: Tensor, mask: Tensor, x0: Tensor, t: Tensor, u: np.ndarray, eps: Tensor, sigma: float,
                mask_mode: str = "per_jet"):
    """losses.py:147-204 with the draws given: x0 [B,N,F] prior sample, t [B], u [B,N] uniforms of np.random.choice,
    eps [B,N,F].  ``vf(t_(B,N), y, mask_eff)``.  Returns (loss, x1 after the in-place re-indexing, mask used, y).
    mask_mode "reference": ``mask_ot`` is the LAST jet's re-indexed mask, shape (N, 1), as the reference leaves it after
    its loop (:189) -- with that mask the reference's own EPiC_encoder raises a shape error (epic.py:370,
    ``z_sum / mask.sum(1)``: (B, H) / (N,)), so the reference's CFM-OT cannot run end to end; this mode only exists to pin
    the coupling and interpolation against the reference's code with a recording stand-in for the network.
    "per_jet": every jet keeps its own re-indexed mask (B, N, 1) -- the evident intent, and what the CUDA path does."""
    x0 = x0.clone()
    x1 = x.clone()
    B, N = x.shape[0], x.shape[1]
    tt = t.unsqueeze(-1).repeat_interleave(N, dim=1).unsqueeze(-1).type_as(x0)            # :154-156
    M = torch.cdist(x0, x1) ** 2                                                          # :163
    masks = []
    for k in range(B):                                                                    # :166-189
        Mk = M[k] / M[k].max()
        pi = emd_uniform(Mk.detach().cpu().numpy())
        p = pi.flatten()
        p = p / p.sum()
        choices = choice_from_uniform(p, u[k])
        i, j = np.divmod(choices, pi.shape[1])
        x0[k] = x0[k, i]
        x1[k] = x1[k, j]
        masks.append(mask[k, j])
    mask_ot = masks[-1] if mask_mode == "reference" else torch.stack(masks)
    mu_t = x0 * tt + x1 * (1 - tt)                                                        # :191
    y = mu_t + sigma * eps                                                                # :193
    ut = (x0 - x1) * mask_ot                                                              # :194-195
    vt = vf(tt.squeeze(-1), y, mask_ot)
    loss = torch.nn.functional.mse_loss(vt, ut, reduction="sum") / mask.sum()             # :203
    return loss, x1, mask_ot, y


# ----------------------------------------------------------------------------------------------
# diffusion
# ----------------------------------------------------------------------------------------------
def _angles(t: Tensor, max_sr: float, min_sr: float):
    start, end = math.acos(max_sr), math.acos(min_sr)                                     # diffusion.py:52-54
    return start + t * (end - start), end - start


def diff_rates(t: Tensor, max_sr: float, min_sr: float):
    a, _ = _angles(t, max_sr, min_sr)
    return torch.cos(a), torch.sin(a)                                                     # :56-58


def diff_betas(t: Tensor, max_sr: float, min_sr: float):
    a, span = _angles(t, max_sr, min_sr)
    return 2 * span * torch.tan(a)                                                        # :67-70


def diffusion_loss(vf: Callable, x: Tensor, mask: Tensor, t: Tensor, z: Tensor, diff_config: Mapping,
                   criterion: str = "huber", mle_loss_weight: float = 0.001) -> Tensor:
    """losses.py:230-285 with (t [B], z [B,N,F]) given.  ``vf(t_(B,N), noisy)``."""
    N = x.shape[1]
    tt = t.unsqueeze(-1).repeat_interleave(N, dim=1).unsqueeze(-1).type_as(x)
    noises = z * mask                                                                     # :240
    times = tt.clone()[:, 0]
    sr, nr = diff_rates(times.view(-1, 1, 1), **diff_config)                              # :255
    noisy = sr * x + nr * noises                                                          # :258
    pred = vf(tt.squeeze(-1), noisy)
    crit = torch.nn.HuberLoss(reduction="none") if criterion == "huber" else torch.nn.MSELoss(reduction="none")
    simple = crit(noises, pred) * mask                                                    # :269
    if mle_loss_weight:
        betas = diff_betas(times.view(-1, 1, 1), **diff_config)
        mle = (betas / nr) * simple
        return simple.sum() / mask.sum() + mle_loss_weight * mle.sum() / mask.sum()       # :276-278
    return simple.sum() / mask.sum()


def ddim_sample(vf: Callable, z: Tensor, n_steps: int, diff_config: Mapping) -> Tensor:
    """solver.py:23-95; ``vf(t_0dim, x)``; returns pred_data of the last step."""
    shape = [-1] + [1] * (z.dim() - 1)
    step = 1 / n_steps
    noisy = z
    tm = torch.ones(z.shape[0])
    nsr, nnr = diff_rates(tm.view(shape), **diff_config)
    pred_data = None
    for _ in range(n_steps):
        sr, nr = nsr, nnr
        pred_noises = vf(tm[0], noisy)
        pred_data = (noisy - nr * pred_noises) / sr
        tm = tm - step
        nsr, nnr = diff_rates(tm.view(shape), **diff_config)
        noisy = nsr * pred_data + nnr * pred_noises
    return pred_data


def em_sample(vf: Callable, z: Tensor, n_steps: int, diff_config: Mapping, noise: Sequence[Tensor]) -> Tensor:
    """solver.py:98-143 with the per-step normal draws given."""
    shape = [-1] + [1] * (z.dim() - 1)
    delta_t = 1 / n_steps
    x_t = z.clone()
    t = torch.ones(z.shape[0])
    for s in range(n_steps):
        pred = vf(t[0], x_t)
        _, nr = diff_rates(t.view(shape), **diff_config)
        sc = -pred / nr
        betas = diff_betas(t.view(shape), **diff_config)
        x_t = x_t + 0.5 * betas * (x_t + 2 * sc) * delta_t
        x_t = x_t + (betas * delta_t).sqrt() * noise[s]
        t = t - delta_t
    return x_t


def pf_ode_sample(vf: Callable, z: Tensor, ode_steps: int, solver: str, diff_config: Mapping) -> Tensor:
    """Probability-flow ODE: dx/dt = -0.5 beta (x - eps_theta / noise_rate)  (flow_matching_module.py:62-69)."""
    def f(t, x):
        shape = [-1] + [1] * (x.dim() - 1)
        _, nr = diff_rates(t.view(shape), **diff_config)
        betas = diff_betas(t.view(shape), **diff_config)
        return -0.5 * betas * (x - vf(t, x) / nr)
    return ode_oracle.integrate(f, z, ode_steps, solver)


# ----------------------------------------------------------------------------------------------
# jet-feature flow
# ----------------------------------------------------------------------------------------------
MLP_BLOCKS = [("mlp1", [64, 64, 64]), ("mlp2", [256, 256, 256]), ("mlp3", [256, 256, 256]), ("mlp4", [64, 64, None])]


def mlp_flow_shapes(features: int, freqs: int, dim_cond: int = 1):
    """[(state_dict prefix, out, in)] of small_cond_MLP_model's linears (mlp.py:34-56)."""
    T = 2 * freqs
    shapes, prev = [], features
    for name, widths in MLP_BLOCKS:
        inp = prev + T + dim_cond
        for li, w in enumerate(widths):
            w = features if w is None else w
            shapes.append((f"{name}.{2 * li}", w, inp))
            inp = w
        prev = inp
    return shapes


def synth_mlp_state_dict(features: int, freqs: int, seed: int):
    """nn.Linear-like init from numpy's legacy RandomState (regenerated in the tests from the seed)."""
    rs = np.random.RandomState(seed)
    sd = {}
    for name, o, i in mlp_flow_shapes(features, freqs):
        k = 1.0 / math.sqrt(i)
        sd[f"{name}.weight"] = torch.from_numpy(rs.uniform(-k, k, size=(o, i)).astype("float32"))
        sd[f"{name}.bias"] = torch.from_numpy(rs.uniform(-k, k, size=(o,)).astype("float32"))
    return sd


def _act(name: str):
    return getattr(torch.nn, name)()


def mlp_flow_forward(sd: Mapping[str, Tensor], t: Tensor, x: Tensor, cond: Tensor, freqs: int, activation: str = "ELU") -> Tensor:
    """CNF.forward of flow_matching_no_sets.py:55-66 + small_cond_MLP_model.forward mlp.py:58-68."""
    fr = torch.arange(1, freqs + 1) * torch.pi
    te = fr * t[..., None]
    te = torch.cat((te.cos(), te.sin()), dim=-1)
    te = te.expand(*x.shape[:-1], -1)
    act = _act(activation)
    h = x
    for name, widths in MLP_BLOCKS:
        h = torch.cat([te, h, cond], dim=-1)
        for li in range(len(widths)):
            h = torch.nn.functional.linear(h, sd[f"{name}.{2 * li}.weight"], sd[f"{name}.{2 * li}.bias"])
            if li + 1 < len(widths):
                h = act(h)
    return h


def mlp_flow_sample(sd, z: Tensor, cond: Tensor, freqs: int, activation: str, ode_steps: int, solver: str = "midpoint") -> Tensor:
    """CNF.decode, flow_matching_no_sets.py:74-93 (midpoint on linspace(1, 0, ode_steps))."""
    return ode_oracle.integrate(lambda t, x: mlp_flow_forward(sd, t, x, cond, freqs, activation), z, ode_steps, solver)


def mlp_flow_loss(sd, x: Tensor, cond: Tensor, t: Tensor, z: Tensor, freqs: int, activation: str, sigma: float) -> Tensor:
    """FlowMatchingLoss.forward on rows (losses.py:38-77, the 2-D branch :49-50): t [B,1], z [B,F]."""
    mask = torch.ones_like(x[..., 0]).unsqueeze(-1)
    y = (1 - t) * x + (sigma + (1 - sigma) * t) * z
    u = ((1 - sigma) * z - x) * mask
    v = mlp_flow_forward(sd, t.squeeze(-1), y, cond, freqs, activation)
    return (v - u).square().sum() / mask.sum()
