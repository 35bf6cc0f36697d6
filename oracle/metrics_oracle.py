"""Oracle (test infrastructure): jet-level observables and Wasserstein metrics used for distribution-level parity.

The reference evaluates generated jets with ``jetnet.evaluation.w1m / w1p`` (third party, unpinned,
particle_fm/data/components/metrics.py:4,122-139): W1 distance of the jet-mass distribution and of the per-particle
feature distributions between random subsamples, mean +- std over ``num_batches`` draws.  jetnet is not installed
here, so this restates the published definition with scipy -- PARITY UNPINNED for the metric itself; it is applied
identically to both sides of every comparison.  Jet mass from massless constituents follows the reference's own
torch formula (data/components/utils.py:65-105), with JetNet's feature order (eta_rel, phi_rel, pt_rel).
The batching scheme is the reference's ``wasserstein_distance_batched`` (metrics.py:11-34).
"""
from __future__ import annotations

import numpy as np
from scipy.stats import wasserstein_distance


def jet_masses(x: np.ndarray) -> np.ndarray:
    """x (B, N, 3) = (eta, phi, pt) of massless particles (padded particles have pt = 0)."""
    eta, phi, pt = x[..., 0], x[..., 1], x[..., 2]
    e, px, py, pz = pt * np.cosh(eta), pt * np.cos(phi), pt * np.sin(phi), pt * np.sinh(eta)
    m2 = e.sum(1) ** 2 - px.sum(1) ** 2 - py.sum(1) ** 2 - pz.sum(1) ** 2
    return np.sign(m2) * np.sqrt(np.abs(m2))


def w1_batched(a: np.ndarray, b: np.ndarray, num_eval_samples: int, num_batches: int, rng) -> tuple:
    w = []
    for _ in range(num_batches):
        w.append(wasserstein_distance(a[rng.choice(len(a), size=num_eval_samples)], b[rng.choice(len(b), size=num_eval_samples)]))
    return float(np.mean(w)), float(np.std(w))


def w1m(x1: np.ndarray, x2: np.ndarray, num_eval_samples: int = 256, num_batches: int = 5, seed: int = 0) -> tuple:
    return w1_batched(jet_masses(x1), jet_masses(x2), num_eval_samples, num_batches, np.random.default_rng(seed))


def w1p(x1: np.ndarray, m1: np.ndarray, x2: np.ndarray, m2: np.ndarray, num_eval_samples: int = 256, num_batches: int = 5,
        seed: int = 0) -> tuple:
    """Mean over the particle features of the W1 distance between the real particles of the two samples."""
    rng = np.random.default_rng(seed)
    means, stds = [], []
    for f in range(x1.shape[-1]):
        a, b = x1[..., f][m1.astype(bool)], x2[..., f][m2.astype(bool)]
        mu, sd = w1_batched(a, b, num_eval_samples * 10, num_batches, rng)
        means.append(mu)
        stds.append(sd)
    return float(np.mean(means)), float(np.mean(stds))
