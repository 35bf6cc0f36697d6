"""Shared helpers of the test-suite: golden fixtures, synthetic weights, module builders."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle import epic_oracle as eo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["c1_jetnet30", "c2_jetnet150", "bare_sincos", "cond_lhco_like", "cond_jetclass_like", "tglobal_plain"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.meta = json.loads(str(z["meta"]))
        self.arr = {k: z[k] for k in z.files if k != "meta"}
        self.ctor = self.meta["ctor"]
        self.cfg = eo.EpicCfg(**self.meta["cfg"])
        self.sd = eo.synth_state_dict(self.cfg, self.meta["wseed"], weight_norm=self.meta["weight_norm"])

    def t(self, key):
        return torch.from_numpy(self.arr[key])

    @property
    def x(self):
        return self.t("x")

    @property
    def mask(self):
        return self.t("mask")

    @property
    def cond(self):
        return self.t("cond") if "cond" in self.arr else None

    def oracle_kwargs(self):
        c = self.ctor
        return dict(t_emb=c["t_emb"], frequencies=c["frequencies"], add_time_to_input=c["add_time_to_input"])

    def oracle_vf(self, sd=None, cond="__own__", mask="__own__"):
        sd = self.sd if sd is None else sd
        cond = self.cond if isinstance(cond, str) else cond
        mask = self.mask if isinstance(mask, str) else mask
        kw = self.oracle_kwargs()
        return lambda t, y: eo.cnf_forward(sd, self.cfg, t, y, cond, mask, **kw)


def full_state_dict(module, sd):
    """Reference-style state_dict (both 'flows.0.' and 'loss.flows.0.' prefixes) from net-level weights."""
    full = {}
    for pre in ("flows.0.", "loss.flows.0."):
        for k, v in sd.items():
            full[pre + "net." + k] = v.clone()
        full[pre + "frequencies"] = module.flows[0].frequencies.detach().clone().cpu()
    return full


def build_module(ctor, sd, loss_type="FM-OT", sigma=1e-4, device=None):
    from particle_fm_b200.models.flow_matching_module import SetFlowMatchingLitModule
    m = SetFlowMatchingLitModule(optimizer=None, loss_type=loss_type, sigma=sigma, **ctor)
    m.load_state_dict(full_state_dict(m, sd), strict=True)
    if device is not None:
        m = m.to(device)
    return m


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
