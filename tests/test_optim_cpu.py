"""CPU tests for the optimizer half of SURVEY.md 8f-1: the restated AdamW arithmetic of oracle/optim_oracle.py against
torch.optim.AdamW + clip_grad_norm_ themselves (what the reference runs), and the host-side contract of
`dyffusion_b200.optim.AdamW` that needs no device (argument validation like torch's, no CPU path)."""
import pytest
import torch

from oracle import optim_oracle as OO
from oracle.synth import synth_tensor
from tests import helpers as H

HYPER = dict(lr=3e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-4)  # experiment/navier_stokes.yaml + optimizer/adamw.yaml
SHAPES = [(64, 9, 3, 3), (64,), (3, 64, 1, 1), (7,), (1,), (128, 130)]


def close(a, b):
    """Equal to 1e-5 of the tensor's scale (moments pass through 0, where a relative bound on single elements is
    meaningless); fp32 arithmetic in torch's order, the rest is FMA contraction / 1-ulp level."""
    b = b.to(a.device)
    return torch.allclose(a, b, rtol=1e-5, atol=1e-5 * float(b.abs().max()))


def update_close(p, want, p0):
    """The accumulated update itself: 1e-4 of its norm, plus two fp32 ulps of the parameter it is added to."""
    err, upd = (p - want).double().norm(), (want - p0).double().norm()
    return float(err) <= 1e-4 * float(upd) + 2.4e-7 * float(want.double().norm())


def case(steps=5, scale=1.0):
    params = [0.1 * synth_tensor(f"opt.p{i}", s) for i, s in enumerate(SHAPES)]
    grads = [[scale * synth_tensor(f"opt.g{k}.{i}", s) for i, s in enumerate(SHAPES)] for k in range(steps)]
    return params, grads


@pytest.mark.parametrize("max_norm", [None, 1.0])
def test_restated_arithmetic_equals_torch_adamw(max_norm):
    params, grads = case()
    want_p, want_m, want_v, norms, _ = OO.reference_steps(params, grads, max_grad_norm=max_norm, **HYPER)
    p, m, v = list(params), [torch.zeros_like(x) for x in params], [torch.zeros_like(x) for x in params]
    for k, gs in enumerate(grads):
        coef = OO.clip_coefficient(gs, max_norm) if max_norm else 1.0
        for i in range(len(p)):
            p[i], m[i], v[i] = OO.restated_step(p[i], gs[i], m[i], v[i], k + 1, clip_coef=coef, **HYPER)
    for i in range(len(p)):
        assert close(p[i], want_p[i]) and close(m[i], want_m[i]) and close(v[i], want_v[i]), i
        assert update_close(p[i], want_p[i], params[i])  # the update itself, not just the parameter
    if max_norm:
        assert all(n > max_norm for n in norms)  # the synthetic gradients are large enough for the clip to be active


def test_host_contract_without_a_device():
    import dyffusion_b200.engine as E
    from dyffusion_b200.optim import AdamW
    p = [torch.nn.Parameter(torch.zeros(4))]
    for bad in (dict(lr=-1.0), dict(eps=-1e-8), dict(betas=(1.0, 0.9)), dict(betas=(0.9, -0.1)), dict(weight_decay=-1.0)):
        with pytest.raises(ValueError):
            AdamW(p, **bad)
        with pytest.raises(ValueError):
            torch.optim.AdamW(p, **bad)  # same errors as the optimizer it replaces
    with pytest.raises(NotImplementedError):
        AdamW(p, amsgrad=True)
    with pytest.raises(E.EngineError):
        AdamW(p)  # CPU parameters: no fallback
    assert E.LIB.dyf_adamw_step(None, None, None, None, 0, 1e-3, 0.9, 0.99, 1e-8, 0.0, 1, 0.0, None, 0, None) == -1
