"""TEST INFRASTRUCTURE -- deterministic synthetic weights, inputs and dropout masks shared by the golden-vector
generator (run against the real reference in the build container) and the parity tests / bench (run anywhere).

Everything is a pure function of (name, shape, seed) through CPU torch generators, so the GPU box regenerates
bit-identical tensors without shipping ~80 MB of weights.
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Optional, Sequence

import torch
from torch import Tensor


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0) -> Dict[str, Tensor]:
    """Fan-in scaled random weights (activations stay O(1) through ~15 layers) and *non-trivial* norm statistics
    (default BatchNorm running stats of mean 0 / var 1 would hide folding bugs)."""
    sd: Dict[str, Tensor] = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        g = _gen(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros((), dtype=torch.long)
        elif leaf == "running_mean":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "g":  # channel-LayerNorm gain
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":  # BatchNorm / GroupNorm gamma
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:  # biases (conv, linear, norm beta)
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        else:  # conv [O, I, kh, kw] / conv-transpose [I, O, kh, kw] / linear [O, I]
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if key.startswith("readout."):  # ConvTranspose2d k4 s2: each output pixel sees I * 2 * 2 taps
                fan_in = shape[0] * 4
            gain = {"final_conv.weight": 0.15, "head.weight": 0.5}.get(key, 1.3)  # keep |net(x)| ~ |x| so that
            sd[key] = (gain / fan_in ** 0.5) * torch.randn(shape, generator=g)   # chained sampler calls stay O(1)
    return sd


def synth_tensor(name: str, shape: Sequence[int], seed: int = 0, kind: str = "normal") -> Tensor:
    g = _gen(name, seed)
    if kind == "normal":
        return torch.randn(tuple(shape), generator=g)
    if kind == "mask":  # Bernoulli(0.1) {0,1} static-condition mask
        return (torch.rand(tuple(shape), generator=g) < 0.1).float()
    raise ValueError(kind)


class SiteDropout:
    """Deterministic stand-in for nn.Dropout used to pin dropout *placement and scaling*: the keep-mask of a
    site is a pure function of (site name, call index at that site, shape, seed)."""

    def __init__(self, seed: int = 0):
        self.seed = seed
        self.calls: Dict[str, int] = {}

    def __call__(self, site: str, x: Tensor, p: float) -> Tensor:
        if p <= 0:
            return x
        n = self.calls.get(site, 0)
        self.calls[site] = n + 1
        keep = torch.rand(x.shape, generator=_gen(f"{site}#{n}", self.seed)) >= p
        return x * keep.to(x.dtype) / (1.0 - p)
