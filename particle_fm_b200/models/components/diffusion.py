"""Cosine variance-preserving diffusion schedule -- mirror of particle_fm/models/components/diffusion.py:9-70.

Host-side scalar math (signal / noise rates and betas of a diffusion time), same formulas and the
same fp32 torch ops as the reference so that the per-step coefficients handed to the CUDA samplers
(``pfm_epic_sample_diffusion``) round like the reference's."""
from __future__ import annotations

import math
from typing import Tuple

import torch as T


class VPDiffusionSchedule:
    def __init__(self, max_sr: float = 1, min_sr: float = 1e-2) -> None:
        self.max_sr = max_sr
        self.min_sr = min_sr

    def __call__(self, time: T.Tensor) -> Tuple[T.Tensor, T.Tensor]:
        return cosine_diffusion_shedule(time, self.max_sr, self.min_sr)

    def get_betas(self, time: T.Tensor) -> T.Tensor:
        return cosine_beta_shedule(time, self.max_sr, self.min_sr)


def _angles(diff_time: T.Tensor, max_sr: float, min_sr: float):
    start_angle = math.acos(max_sr)
    end_angle = math.acos(min_sr)
    return start_angle + diff_time * (end_angle - start_angle), end_angle - start_angle


def cosine_diffusion_shedule(diff_time: T.Tensor, max_sr: float = 1, min_sr: float = 1e-2):
    """(signal_rate, noise_rate) = (cos, sin) of the diffusion angle (diffusion.py:23-58)."""
    ang, _ = _angles(diff_time, max_sr, min_sr)
    return T.cos(ang), T.sin(ang)


def cosine_beta_shedule(diff_time: T.Tensor, max_sr: float = 1, min_sr: float = 1e-2) -> T.Tensor:
    """beta(t) = 2 (end - start) tan(angle) (diffusion.py:61-70)."""
    ang, span = _angles(diff_time, max_sr, min_sr)
    return 2 * span * T.tan(ang)
