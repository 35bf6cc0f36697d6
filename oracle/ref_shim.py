"""Oracle (test infrastructure): load the UNMODIFIED reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  The reference package
cannot be imported as shipped because pytorch_lightning, torchdyn, zuko, POT, hydra ... are not
installed; this registers minimal stand-ins for exactly those third-party imports and
then imports ``particle_fm.models.flow_matching_module`` from the read-only tree, so CNF,
SetFlowMatchingLitModule, the losses and EPiC_encoder run with the reference's own code.

The ``torchdyn`` stand-in below is a transcription of torchdyn 1.0.x's fixed-step driver in its
original object form (solver classes with ``step``; ``_fixed_odeint`` loop) -- it is NOT the real
package, which is why the integrator stays "parity unpinned"; it exists so that the reference's
``sample()`` / ``decode()`` plumbing around the integrator can be exercised unmodified and to
cross-check oracle/ode_oracle.py's closed-form time grid against the loop form.
"""
from __future__ import annotations

import importlib
import inspect
import logging
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("PFM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "particle_fm", "models", "components"))


class _LightningModuleStandIn(nn.Module):
    """What the reference module needs from pl.LightningModule: hparams, device, log()."""

    def save_hyperparameters(self, *a, **k):
        frame = inspect.currentframe().f_back
        args, _, _, values = inspect.getargvalues(frame)
        hp = types.SimpleNamespace(**{n: values[n] for n in args if n != "self"})
        hp.__dict__.update({})
        self.hparams = hp

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass

    current_epoch = 1


# ---- torchdyn stand-in (transcribed driver, see module docstring) -------------------------------
class _Euler:
    def step(self, f, x, t, dt, k1=None, args=None):
        if k1 is None:
            k1 = f(t, x)
        return None, x + dt * k1, None


class _Midpoint:
    def step(self, f, x, t, dt, k1=None, args=None):
        if k1 is None:
            k1 = f(t, x)
        x_mid = x + 0.5 * dt * k1
        return None, x + dt * f(t + 0.5 * dt, x_mid), None


def _fixed_odeint(f, x, t_span, solver):
    t, dt = t_span[0], t_span[1] - t_span[0]
    sol = [x]
    steps = 1
    while steps <= len(t_span) - 1:
        _, x, _ = solver.step(f, x, t, dt)
        t = t + dt
        sol.append(x)
        if steps < len(t_span) - 1:
            dt = t_span[steps + 1] - t
        steps += 1
    return torch.stack(sol)


class _NeuralODE(nn.Module):
    def __init__(self, vector_field, solver="euler", sensitivity="adjoint", **kw):
        super().__init__()
        self.vf = vector_field
        if solver not in ("euler", "midpoint"):
            raise NotImplementedError(f"torchdyn stand-in: solver {solver}")
        self.solver = _Euler() if solver == "euler" else _Midpoint()

    def trajectory(self, x, t_span):
        t_span = t_span.to(x)
        f = self.vf
        if t_span[1] < t_span[0]:
            f_ = lambda t, x: -f(-t, x)
            t_span = -t_span
        else:
            f_ = f
        return _fixed_odeint(f_, x, t_span, self.solver)


def _unavailable(name):
    def fn(*a, **k):
        raise RuntimeError(f"{name} is a stand-in: the real third-party package is not installed")
    return fn


_loaded = None


def load():
    """Returns a namespace with the reference's modules: fm (flow_matching_module), epic, losses,
    time_emb, droid."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    pkg = mod("particle_fm")
    pkg.__path__ = [os.path.join(REF_ROOT, "particle_fm")]
    utl = mod("particle_fm.utils")
    utl.__path__ = []
    mod("particle_fm.utils.pylogger", get_pylogger=logging.getLogger)
    pl = mod("pytorch_lightning", LightningModule=_LightningModuleStandIn)
    pl.__path__ = []
    td = mod("torchdyn")
    td.__path__ = []
    mod("torchdyn.core", NeuralODE=_NeuralODE)
    zk = mod("zuko")
    zk.__path__ = []
    mod("zuko.utils", odeint=_unavailable("zuko.utils.odeint"))
    mod("ot", emd=_unavailable("ot.emd"), unif=_unavailable("ot.unif"))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fm = importlib.import_module("particle_fm.models.flow_matching_module")
    ns = types.SimpleNamespace(
        fm=fm,
        epic=sys.modules["particle_fm.models.components.epic"],
        losses=sys.modules["particle_fm.models.components.losses"],
        time_emb=sys.modules["particle_fm.models.components.time_emb"],
        droid=sys.modules["particle_fm.models.components.droid_transformer"],
    )
    _loaded = ns
    return ns
