"""Oracle (test infrastructure): fixed-step ODE integration as the reference performs it.

PARITY UNPINNED for this file: the arithmetic is third-party ``torchdyn`` (unpinned,
requirements.txt:25; call sites flow_matching_module.py:278-287 ``NeuralODE(..., solver="euler" |
"midpoint").trajectory(z, linspace(1, 0, ode_steps))[-1]``).  torchdyn is not installed here and
not vendored, so this restates torchdyn 1.0.x's published algorithm:

  torchdyn/numerics/odeint.py::odeint            reversed time  (t_span[1] < t_span[0]):
                                                 f_(t, x) = -f(-t, x);  t_span = -t_span
  torchdyn/numerics/odeint.py::_fixed_odeint     t = t_span[0]; dt = t_span[1]-t_span[0];
                                                 loop: x = step(f_, x, t, dt); t = t + dt;
                                                       if more steps: dt = t_span[k+1] - t
  torchdyn/numerics/solvers/ode.py::Euler.step   x + dt * k1,                 k1 = f_(t, x)
                                  ::Midpoint.step  x + dt * f_(t + 0.5*dt, x + 0.5*dt*k1)

``ode_steps`` is the number of GRID POINTS: ode_steps-1 steps; midpoint does 2 evaluations a step.
All of t, dt are fp32 scalars; the model is evaluated at ``-t`` of the reversed grid.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch

Tensor = torch.Tensor


def time_grid(ode_steps: int, solver: str) -> Tuple[Tensor, Tensor]:
    """The fp32 times the model is evaluated at, and the fp32 dt of every step, exactly as the
    recurrence above produces them.  Returns (t_eval [n_nfe], dt [ode_steps-1])."""
    t_span = -torch.linspace(1.0, 0.0, ode_steps)      # flow_matching_module.py:280,285 + reversal
    t = t_span[0]
    dt = t_span[1] - t_span[0]
    t_eval: List[Tensor] = []
    dts: List[Tensor] = []
    n = ode_steps - 1
    for k in range(1, n + 1):
        t_eval.append(-t)                              # f_(t, x) = -f(-t, x)
        if solver == "midpoint":
            t_eval.append(-(t + 0.5 * dt))
        elif solver != "euler":
            raise NotImplementedError(solver)
        dts.append(dt)
        t = t + dt
        if k < n:
            dt = t_span[k + 1] - t
    return torch.stack(t_eval), torch.stack(dts)


def integrate(f: Callable[[Tensor, Tensor], Tensor], x: Tensor, ode_steps: int, solver: str,
              return_evals: bool = False):
    """Integrate dx/dt = f(t, x) from t=1 to t=0 on linspace(1, 0, ode_steps) (reference decode()).
    ``f`` receives a 0-dim fp32 time tensor like torchdyn hands the wrapped CNF.
    With return_evals=True also returns the list of (t, x_in, v) of every network evaluation
    (used for teacher-forced per-step parity)."""
    t_eval, dts = time_grid(ode_steps, solver)
    evals = []
    j = 0
    for k in range(ode_steps - 1):
        dt = dts[k]
        v = f(t_eval[j], x)
        if return_evals:
            evals.append((t_eval[j], x, v))
        k1 = -v
        j += 1
        if solver == "euler":
            x = x + dt * k1
        else:
            x_mid = x + 0.5 * dt * k1
            v2 = f(t_eval[j], x_mid)
            if return_evals:
                evals.append((t_eval[j], x_mid, v2))
            j += 1
            x = x + dt * (-v2)
    return (x, evals) if return_evals else x
