"""GPU: distribution-level parity (north star: "jet-mass / pT W1 metrics within the reference's seed-to-seed spread").
Same weights and masks; the CUDA samples (fp32 and bf16 paths, same noise as oracle run A) are compared with the
oracle's samples through W1m / W1p, against the W1 between two ORACLE runs that differ only in the noise seed."""
import numpy as np
import pytest
import torch

from oracle import epic_oracle as eo
from oracle import loss_oracle as lo
from oracle import metrics_oracle as mo

from helpers import Golden, build_module

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_w1_metrics_within_seed_to_seed_spread(lib_built):
    g = Golden("c1_jetnet30")
    B, N, steps = 768, 30, 12
    _, mask, _ = eo.synth_cloud(B, N, 3, 2718)
    vf = g.oracle_vf(mask=mask)

    def oracle_run(seed):
        torch.manual_seed(seed)
        z = torch.randn(B, N, 3)
        with torch.no_grad():
            return z, lo.sample(vf, z, mask, "midpoint", steps).numpy()

    zA, oA = oracle_run(1)
    spreads_m, spreads_p = [], []
    mk = mask.squeeze(-1).numpy()
    for seed in (2, 3, 4, 5, 6):                      # seed-to-seed spread of the reference path itself
        _, oS = oracle_run(seed)
        spreads_m.append(mo.w1m(oA, oS)[0])
        spreads_p.append(mo.w1p(oA, mk, oS, mk)[0])
    spread_m, spread_p = float(np.mean(spreads_m)), float(np.mean(spreads_p))
    m = build_module(g.ctor, g.sd, device=DEV)
    for prec in ("fp32", "bf16"):
        m.set_precision(prec)
        with torch.no_grad():
            s = m.flows[0].decode((zA * mask).to(DEV), None, mask.to(DEV), "midpoint", steps).cpu().numpy()
        wm, wp = mo.w1m(oA, s)[0], mo.w1p(oA, mk, s, mk)[0]
        # same noise: the sampling floor of the batched estimator (two random subsamples of the SAME distribution)
        floor_m, floor_p = mo.w1m(oA, oA)[0], mo.w1p(oA, mk, oA, mk)[0]
        print(f"{prec}: W1m {wm:.3e} (same-sample floor {floor_m:.3e}, seed-to-seed {spread_m:.3e}); "
              f"W1p {wp:.3e} (floor {floor_p:.3e}, seed-to-seed {spread_p:.3e})")
        assert wm <= max(spread_m, 1.5 * floor_m) and wp <= max(spread_p, 1.5 * floor_p)
        # and directly: the jet masses agree jet by jet far below the spread
        dm = np.abs(mo.jet_masses(s) - mo.jet_masses(oA)).mean()
        assert dm < (2e-2 if prec == "bf16" else 1e-4) * max(spread_m, 1e-9) * 50 + 1e-6
