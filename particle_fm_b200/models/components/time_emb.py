"""Time codes, host side.  Mirrors particle_fm/models/components/time_emb.py (cosine_encoding :49-96,
CosineEncoding :25-46) and CNF.time_embedding's "sincos" branch (flow_matching_module.py:208-211).

The cosine code is chaotic in t (frequencies up to e^31), so the op ORDER below is the
reference's: ((t + min) * freqs) * pi / (max + min), all fp32.  The tables are tiny ([n_evals, 32]);
they are evaluated with torch ops and handed to the CUDA kernels."""
from __future__ import annotations

import math

import torch

Tensor = torch.Tensor


_EXP_FREQS = {}


def cosine_encoding(x: Tensor, outp_dim: int = 32, min_value: float = 0.0, max_value: float = 1.0,
                    frequency_scaling: str = "exponential") -> Tensor:
    if x.shape[-1] != 1 or x.dim() == 1:
        x = x.unsqueeze(-1)
    if frequency_scaling == "exponential":
        # exp() is evaluated on the HOST and the 32-entry table moved: CPU and CUDA expf differ in the last
        # bit for some k, and one ulp of e^31 turns the high channels into a different hash of t.  The host
        # table is what the CPU reference (and the oracle) use; see DESIGN.md "time code".
        key = (outp_dim, x.device)
        freqs = _EXP_FREQS.get(key)
        if freqs is None:              # cached per device: no host -> device copy per call (illegal inside CUDA-graph capture)
            freqs = torch.arange(outp_dim).exp().to(x.device)
            _EXP_FREQS[key] = freqs
    elif frequency_scaling == "linear":
        freqs = torch.arange(1, outp_dim + 1, device=x.device)
    else:
        raise RuntimeError(f"Unrecognised frequency scaling: {frequency_scaling}")
    return torch.cos((x + min_value) * freqs * math.pi / (max_value + min_value))


class CosineEncoding:
    def __init__(self, outp_dim: int = 32, min_value: float = 0.0, max_value: float = 1.0,
                 frequency_scaling: str = "exponential") -> None:
        self.outp_dim = outp_dim
        self.min_value = min_value
        self.max_value = max_value
        self.frequency_scaling = frequency_scaling

    def __call__(self, inpt: Tensor) -> Tensor:
        return cosine_encoding(inpt, self.outp_dim, self.min_value, self.max_value, self.frequency_scaling)


def sincos_encoding(t: Tensor, frequencies: Tensor) -> Tensor:
    """cat(cos(f t), sin(f t)) with f = 2**k * pi (flow_matching_module.py:172, :208-210)."""
    a = frequencies * t[..., None]
    return torch.cat((a.cos(), a.sin()), dim=-1)
