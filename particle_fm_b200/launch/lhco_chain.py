"""LHCO two-step generation chained on the device.

Reference workflow: the jet-feature flow (``FLowMatchingNoSetsLitModule``, experiment/lhco/jet_features.yaml) samples
the dijet features conditioned on m_jj, a script writes them to a conditioning file, and
``scripts/generate_data_lhco.py:143-176`` reads that file back, normalises the jet features and calls
``generate_data`` once per jet with the EPiC particle model (experiment/lhco/both_jets.yaml).  Here step 1's result
never leaves the GPU: un-normalisation, the multiplicity -> mask conversion and the normalisation of the step-2
conditioning are device ops on the sampler's output, and step 2 consumes them directly."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from ..utils.data_generation import generate_data


def _affine(t: torch.Tensor, mean, std, sigma: float, inverse: bool) -> torch.Tensor:
    mean = torch.as_tensor(mean, dtype=torch.float32, device=t.device)
    std = torch.as_tensor(std, dtype=torch.float32, device=t.device)
    if inverse:                                    # data/components/utils.py:183-200
        return t * (std / sigma) + mean
    return (t - mean) / (std / sigma)              # normalize_tensor, :164-180


@torch.no_grad()
def generate_lhco_chain(jet_model, particle_model, mjj: torch.Tensor, *, jet_means: Sequence[float], jet_stds: Sequence[float],
                        mjj_mean: float, mjj_std: float, cond_means: Sequence[float], cond_stds: Sequence[float],
                        normalize_sigma: float = 5, batch_size: int = 1024, ode_solver: str = "midpoint",
                        jet_ode_steps: int = 100, ode_steps: int = 100, particle_means=None, particle_stds=None,
                        log_pt: bool = False, pt_standardization: bool = False, device: str = "cuda",
                        jets: Sequence[int] = (0, 1)) -> Dict[str, object]:
    """m_jj [n] (raw) -> jet features of both jets (step 1) -> particle clouds of both jets (step 2).

    Returns {"jet_features": (n, 2, 5) device tensor (pt, eta, phi, m, n_particles, raw units),
             "mask": (n, 2, N, 1) device tensor, "particle_data": np.ndarray (n, 2, N, F), "generation_time": seconds}."""
    dev = torch.device(device)
    n = int(mjj.shape[0])
    N = int(particle_model.hparams.num_particles)
    jet_model = jet_model.to(dev)
    # ---- step 1: jet features | m_jj   (flow_matching_no_sets.py:212-238)
    cond1 = _affine(mjj.to(dev, torch.float32).reshape(n, 1), [mjj_mean], [mjj_std], normalize_sigma, inverse=False)
    feats = jet_model.sample(n, cond=cond1, ode_solver=ode_solver, ode_steps=jet_ode_steps)          # (n, 10) normalised
    feats = _affine(feats, jet_means, jet_stds, normalize_sigma, inverse=True).reshape(n, 2, -1)     # (n, 2, 5) raw
    mult = feats[..., 4].round().clamp(1, N).to(torch.int64)                                          # particle multiplicity
    feats = torch.cat([feats[..., :4], mult.unsqueeze(-1).to(feats.dtype)], dim=-1)
    mask = (torch.arange(N, device=dev).view(1, 1, N) < mult.unsqueeze(-1)).to(torch.float32).unsqueeze(-1)
    # ---- step 2: particles | jet features, once per jet   (scripts/generate_data_lhco.py:129-176)
    parts, t_total = [], 0.0
    for j in jets:
        cond2 = _affine(feats[:, j, :len(cond_means)], cond_means, cond_stds, normalize_sigma, inverse=False)
        data, t = generate_data(particle_model, num_jet_samples=n, batch_size=batch_size, cond=cond2, device=device,
                                variable_set_sizes=True, mask=mask[:, j], normalized_data=particle_means is not None,
                                normalize_sigma=normalize_sigma, means=particle_means, stds=particle_stds, log_pt=log_pt,
                                pt_standardization=pt_standardization, verbose=False, ode_solver=ode_solver,
                                ode_steps=ode_steps)
        parts.append(data)
        t_total += t
    import numpy as np
    return {"jet_features": feats, "mask": mask, "particle_data": np.stack(parts, axis=1), "generation_time": t_total}
