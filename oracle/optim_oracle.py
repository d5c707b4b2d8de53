"""TEST INFRASTRUCTURE ONLY -- the checker for `dyffusion_b200.optim` / `dyf_adamw_step` (optimizer half).

The reference's optimizer step is third-party code: `torch.optim.AdamW` (built by `BaseExperiment._get_optim`,
`src/experiment_types/_base_experiment.py:711-725`, torch>=1.8 un-pinned in setup.py:90) after Lightning's
`gradient_clip_val` (`torch.nn.utils.clip_grad_norm_`).  Both are importable wherever torch is, so the oracle CALLS them on
CPU (`reference_steps`) -- the parity is pinned against the implementation the reference itself runs, in the installed
torch 2.11 -- and additionally restates the arithmetic (`restated_step`, torch/optim/adamw.py::_single_tensor_adamw +
torch/nn/utils/clip_grad.py) so that the two can be compared on CPU (tests/test_optim_cpu.py).  Only tests/ may import this."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch


def reference_steps(params: Sequence[torch.Tensor], grads_per_step: Sequence[Sequence[torch.Tensor]], *, lr, betas, eps,
                    weight_decay, max_grad_norm: Optional[float]):
    """-> (final params, exp_avg, exp_avg_sq, [total grad norm per step]) after running torch's own AdamW on CPU."""
    ps = [torch.nn.Parameter(p.clone()) for p in params]
    opt = torch.optim.AdamW(ps, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
    norms = []
    for grads in grads_per_step:
        for p, g in zip(ps, grads):
            p.grad = g.clone()
        if max_grad_norm:
            norms.append(float(torch.nn.utils.clip_grad_norm_(ps, max_grad_norm)))
        opt.step()
    return ([p.detach() for p in ps], [opt.state[p]["exp_avg"] for p in ps], [opt.state[p]["exp_avg_sq"] for p in ps], norms,
            opt.state_dict())


def restated_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, *, lr, betas, eps, weight_decay,
                  clip_coef: float = 1.0):
    """One AdamW update of one tensor in fp32, in the order of torch's single-tensor implementation."""
    b1, b2 = betas
    g = g * clip_coef
    p = p * (1 - lr * weight_decay)
    m = m + (g - m) * (1 - b1)
    v = v * b2 + (1 - b2) * g * g
    step_size = lr / (1 - b1 ** step)
    denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
    return p - step_size * (m / denom), m, v


def clip_coefficient(grads: Sequence[torch.Tensor], max_norm: float) -> float:
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    return float(torch.clamp(max_norm / (total + 1e-6), max=1.0))
