"""Batched generation driver -- drop-in for particle_fm/utils/data_generation.py::generate_data (:17-174).

Same signature, same batching / warm-up / timing rule (the clock starts with the second batch and
stops before the remainder batch, :82-83,125,173) and the same post-processing (inverse
normalisation, log-pt, masking), so eval callbacks and scripts can call it unchanged.  Each batch
is one fused CUDA launch through ``model.sample``.

On a CUDA device the post-processing of :105-123 runs ON THE DEVICE (``pfm_postprocess``): one kernel per batch
applies inverse normalisation / log-pt / masking and writes the final values straight into the batch's slice of ONE
pinned host buffer for the whole request -- no per-batch ``.cpu()``, no host-side Python loop over features, no growing
``torch.cat`` (:123), and batch k+1's noise draw / launch overlaps batch k's integration.  The host only synchronises
where the reference's clock is read.
"""
from __future__ import annotations

import time

import numpy as np
import torch


def inverse_normalize_tensor(tensor, mean, std, sigma=5):
    """tensor[..., i] * (std[i] / sigma) + mean[i]  (particle_fm/data/components/utils.py:183-200)."""
    for i in range(len(mean)):
        tensor[..., i] = (tensor[..., i] * (std[i] / sigma)) + mean[i]
    return tensor


def _post_process(batch, mask_batch, normalized_data, normalize_sigma, means, stds, log_pt, pt_standardization,
                  variable_set_sizes):
    if normalized_data:
        if pt_standardization:       # data_generation.py:106-115 (sigma 10 / 5 hard-coded in the reference)
            batch[..., :2] = inverse_normalize_tensor(batch[..., :2], means[:2], stds[:2], sigma=10)
            batch[..., 2] = inverse_normalize_tensor(batch[..., 2], [means[2]], [stds[2]], sigma=5)
        else:
            batch = inverse_normalize_tensor(batch, means, stds, sigma=normalize_sigma)
        if log_pt:
            batch[..., 2] = 1.0 - np.exp(batch[..., 2])
    if variable_set_sizes:
        batch = batch * mask_batch
    return batch


def generate_data(model, num_jet_samples: int, batch_size: int = 256, cond: torch.Tensor = None, device: str = "cuda",
                  variable_set_sizes: bool = False, mask: torch.Tensor = None, normalized_data: bool = False,
                  normalize_sigma: int = 5, means=None, stds=None, log_pt: bool = False,
                  pt_standardization: bool = False, shuffle_mask: bool = False, verbose: bool = True,
                  ode_solver: str = "midpoint", ode_steps: int = 100):
    """Returns (np.ndarray (num_jet_samples, num_particles, features), generation_time seconds)."""
    if variable_set_sizes and mask is None:
        raise ValueError("Please use mask when using variable_set_sizes=True")
    if mask is not None and len(mask) != num_jet_samples:
        raise ValueError(f"Mask should have the same length as num_jet_samples ({len(mask)} != {num_jet_samples})")
    if verbose:
        print(f"Generating data ({num_jet_samples} samples). Device: {torch.device(device)}")
    parts = []
    start_time = 0
    n_full = num_jet_samples // batch_size
    model = model.to(torch.device(device))
    if torch.device(device).type == "cuda" and hasattr(model, "hparams") and hasattr(model.hparams, "num_particles"):
        return _generate_data_device(model, num_jet_samples, batch_size, cond, torch.device(device), variable_set_sizes, mask,
                                     normalized_data, normalize_sigma, means, stds, log_pt, pt_standardization, shuffle_mask,
                                     ode_solver, ode_steps)

    def one_batch(n, cond_batch, mask_batch):
        with torch.no_grad():
            out = model.sample(n_samples=n, cond=cond_batch, mask=mask_batch, ode_solver=ode_solver,
                               ode_steps=ode_steps).cpu()
        return _post_process(out, mask_batch, normalized_data, normalize_sigma, means, stds, log_pt,
                             pt_standardization, variable_set_sizes)

    for i in range(n_full):
        cond_batch = cond[i * batch_size:(i + 1) * batch_size] if cond is not None else None
        if i == 1:
            start_time = time.time()
        if variable_set_sizes:
            if shuffle_mask:
                mask = mask[np.random.permutation(len(mask))]
                mask_batch = mask[:batch_size]
            else:
                mask_batch = mask[i * batch_size:(i + 1) * batch_size]
        else:
            mask_batch = None
        parts.append(one_batch(batch_size, cond_batch, mask_batch))
    end_time = time.time()
    rem = num_jet_samples - n_full * batch_size
    if rem != 0:
        cond_batch = cond[-rem:] if cond is not None else None
        if variable_set_sizes:
            if shuffle_mask:
                mask = mask[np.random.permutation(len(mask))]
            mask_batch = mask[-rem:]
        else:
            mask_batch = None
        parts.append(one_batch(rem, cond_batch, mask_batch))
    particle_data_sampled = np.array(torch.cat(parts)) if parts else np.zeros((0,))
    return particle_data_sampled, end_time - start_time


def post_coefficients(features, normalized_data, normalize_sigma, means, stds, log_pt, pt_standardization):
    """(scale, shift, log_col, first_only_col) of the affine inverse normalisation as python floats, evaluated like the
    reference does: ``std[i] / sigma`` in the caller's own number types, then rounded to fp32 by the multiply
    (data_generation.py:105-117, data/components/utils.py:183-200).  first_only_col = 2 under pt_standardization: the
    reference hands the 2-D slice ``batch[..., 2]`` to inverse_normalize_tensor, whose ``tensor[..., 0]`` then un-normalises
    particle 0 of every jet only (:111-113) -- reproduced as is."""
    if not normalized_data:
        return None, None, -1, -1
    scale, shift = [], []
    for i in range(features):
        if pt_standardization:                    # :106-112: columns 0,1 with sigma=10, column 2 with sigma=5, others untouched
            if i < 2:
                scale.append(float(stds[i] / 10)); shift.append(float(means[i]))
            elif i == 2:
                scale.append(float(stds[2] / 5)); shift.append(float(means[2]))
            else:
                scale.append(1.0); shift.append(0.0)
        elif i < len(means):
            scale.append(float(stds[i] / normalize_sigma)); shift.append(float(means[i]))
        else:
            scale.append(1.0); shift.append(0.0)
    return scale, shift, (2 if log_pt else -1), (2 if pt_standardization else -1)


def _generate_data_device(model, num_jet_samples, batch_size, cond, device, variable_set_sizes, mask, normalized_data,
                          normalize_sigma, means, stds, log_pt, pt_standardization, shuffle_mask, ode_solver, ode_steps):
    from ..engine import postprocess_into
    N, F = int(model.hparams.num_particles), int(model.hparams.features)
    scale, shift, log_col, first_only = post_coefficients(F, normalized_data, normalize_sigma, means, stds, log_pt,
                                                          pt_standardization)
    out = torch.empty((num_jet_samples, N, F), dtype=torch.float32, pin_memory=True)
    n_full = num_jet_samples // batch_size
    start_time = 0

    def one_batch(lo, n, cond_batch, mask_batch):
        with torch.no_grad():
            x = model.sample(n_samples=n, cond=cond_batch, mask=mask_batch, ode_solver=ode_solver, ode_steps=ode_steps)
            postprocess_into(x, mask_batch if variable_set_sizes else None, out[lo:lo + n], scale, shift, log_col, first_only)

    for i in range(n_full):
        cond_batch = cond[i * batch_size:(i + 1) * batch_size] if cond is not None else None
        if i == 1:
            torch.cuda.synchronize(device)         # the reference's clock starts once the first batch is complete (:82-83)
            start_time = time.time()
        if variable_set_sizes:
            if shuffle_mask:
                mask = mask[np.random.permutation(len(mask))]
                mask_batch = mask[:batch_size]
            else:
                mask_batch = mask[i * batch_size:(i + 1) * batch_size]
        else:
            mask_batch = None
        one_batch(i * batch_size, batch_size, cond_batch, mask_batch)
    torch.cuda.synchronize(device)
    end_time = time.time()
    rem = num_jet_samples - n_full * batch_size
    if rem != 0:
        cond_batch = cond[-rem:] if cond is not None else None
        if variable_set_sizes:
            if shuffle_mask:
                mask = mask[np.random.permutation(len(mask))]
            mask_batch = mask[-rem:]
        else:
            mask_batch = None
        one_batch(n_full * batch_size, rem, cond_batch, mask_batch)
        torch.cuda.synchronize(device)
    return out.numpy(), end_time - start_time
