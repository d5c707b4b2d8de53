"""GPU parity of the DYffusion objective's forward half: the drop-in's `p_losses` on the CUDA engine in
validation mode (eval, no_grad -- the reference's `val/loss`) against the reference's own values (tests/golden/p_losses_kat.json)
and the oracle.  Stated tolerance: 2e-2 relative on each loss term (L1 means over bf16-operand forwards chained up to four
deep: interpolator -> forecaster -> interpolator -> forecaster)."""
import json

import pytest
import torch

from tests import helpers as H
from tests import test_training_cpu as T
from tests.gpu_helpers import build_dyffusion

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(T.CASES))
def test_validation_loss_vs_reference_values(name):
    import dyffusion_b200.engine as E
    dataset, horizon, ov, steps = T.CASES[name]
    dyf = build_dyffusion(dataset, horizon=horizon, enable_interpolator_dropout=False, loss_function="l1", **ov)
    last, cond, static = T._inputs(name, dataset, len(steps))
    real, counter = torch.randn_like, {"n": 0}

    def fake(x):
        counter["n"] += 1
        return H.synth_tensor(f"{name}.noise{counter['n'] - 1}", tuple(x.shape)).to(x.device)

    before = E.launch_count()
    torch.randn_like = fake
    try:
        with torch.no_grad():
            got = dyf.p_losses(last.cuda(), cond.cuda(), torch.tensor(steps).cuda(),
                               static_condition=None if static is None else static.cuda())
    finally:
        torch.randn_like = real
    assert E.launch_count() > before
    with open(T.KAT) as f:
        want = json.load(f)[name]
    oracle = T._oracle_losses(name)
    assert counter["n"] == want["noise_draws"]
    for k, gk in (("loss", "loss"), ("loss_forward", "val/loss_forward"), ("loss_forward2", "val/loss_forward2")):
        g = float(got[gk])
        assert abs(g - want[k]) <= 2e-2 * max(abs(want[k]), 1e-3), (name, k, g, want[k])
        assert abs(g - float(oracle[k])) <= 2e-2 * max(abs(float(oracle[k])), 1e-3), (name, k, g, float(oracle[k]))
    print(name, {k: (round(float(got[gk]), 5), round(want[k], 5)) for k, gk in (("loss", "loss"),)})


def test_training_through_the_engine_fails_loudly():
    dyf = build_dyffusion("spring", horizon=4, loss_function="l1")
    dyf.train()
    last, cond, static = T._inputs("spring_h4", "spring", 2)
    with pytest.raises(NotImplementedError):
        dyf.p_losses(last.cuda(), cond.cuda(), torch.tensor([0, 1]).cuda(), static_condition=static.cuda())
