"""Randomised differential test of the rollout's host logic against the reference's `_evaluation_step`
(src/experiment_types/forecasting_multi_horizon.py:115-238) with the network replaced by a cheap deterministic stand-in on
BOTH sides: random horizon / window / ensemble size / batch / autoregressive steps / prediction horizons that are not a
multiple of the horizon, boundary-condition callables that depend on the running time.  Exercises what the fixed cases do
not: windows > 1 (hand-off of several frames, `torch.cat` order), early termination inside an autoregressive step, the time
bookkeeping.  Build container only."""
import random

import numpy as np
import pytest
import torch

from oracle import configs as C

pytestmark = pytest.mark.needs_reference
CH, HW = 4, (10, 10)


class FakeDiffusion(torch.nn.Module):
    """`t{i}_preds` = a fixed function of the LAST frame of the stacked window, the whole window and the static condition."""

    def __init__(self, horizon):
        super().__init__()
        self.horizon, self.num_timesteps = horizon, horizon
        self.hparams = {"timesteps": horizon}
        self.calls = []

    def sample_loop(self, initial_condition, static_condition=None, log_every_t=None, num_predictions=None):
        raise NotImplementedError

    def predict_forward(self, inputs, condition=None, metadata=None, num_predictions=None, **kw):
        self.calls.append((tuple(inputs.shape), num_predictions))
        last = inputs[:, -CH:]
        whole = inputs.reshape(inputs.shape[0], -1, CH, *HW).mean(dim=1)
        s = 0.0 if condition is None else condition.sum(dim=1, keepdim=True)
        return {f"t{i}_preds": torch.tanh(last * (0.5 + 0.1 * i) + 0.3 * whole + 0.01 * s) for i in range(1, self.horizon + 1)}


def _reference_experiment(horizon, window, members, ar_steps, prediction_horizon, noise=0.0):
    from oracle import ref_build
    saved = C.DATASETS["spring"]["datamodule"]["window"]
    C.DATASETS["spring"]["datamodule"]["window"] = window
    try:
        ipol = ref_build.build_interpolator("spring", horizon=horizon)
        exp = ref_build.build_dyffusion("spring", ipol, horizon=horizon)
    finally:
        C.DATASETS["spring"]["datamodule"]["window"] = saved
    exp.hparams.num_predictions = members
    exp.hparams.autoregressive_steps = ar_steps
    exp.hparams.prediction_inputs_noise = noise
    if prediction_horizon is not None:
        exp.datamodule_config["prediction_horizon"] = prediction_horizon
    exp.model = FakeDiffusion(horizon)
    return exp


def test_rollout_agrees_with_the_reference_on_random_configurations():
    from dyffusion_b200.rollout import MultiHorizonRollout
    rng = random.Random(7)
    seen_window, seen_partial = set(), 0
    for trial in range(25):
        horizon, window = rng.randint(2, 5), rng.randint(1, 3)
        window = min(window, horizon)
        members, batch = rng.randint(1, 3), rng.randint(1, 3)
        if rng.random() < 0.5:
            ar_steps, pred_h = rng.randint(0, 2), None
            total = horizon * (ar_steps + 1)
        else:
            ar_steps, pred_h = 0, rng.randint(1, 3 * horizon)
            total = pred_h
            seen_partial += int(pred_h % horizon != 0)
        seen_window.add(window)
        g = torch.Generator().manual_seed(trial)
        batch_d = {"dynamics": torch.randn(batch, window + total + rng.randint(0, 2), CH, *HW, generator=g),
                   "condition": (torch.rand(batch, 1, *HW, generator=g) < 0.3).float(), "metadata": {"tag": torch.arange(batch)}}
        t0, dt = (0.25, 0.5) if rng.random() < 0.5 else (torch.rand(batch, generator=g), torch.full((batch,), 0.1))
        times = {"ref": [], "mine": []}

        def make_bc(log):
            def bc(preds, targets, metadata, time):
                log.append(time if isinstance(time, float) else time.clone())
                shift = time if isinstance(time, float) else time.view(-1, *[1] * (targets.ndim - 1))
                preds += 0.01 * shift + 0.001 * targets  # in place, broadcast over a leading ensemble axis
                return preds
            return bc

        noise = rng.choice([0.0, 0.0, 0.1])  # `prediction_inputs_noise`: per-member input noise from the torch RNG (:523-528)
        exp = _reference_experiment(horizon, window, members, ar_steps, pred_h, noise)
        torch.manual_seed(100 + trial)
        want = exp._evaluation_step({k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch_d.items()}, 0, "test",
                                    boundary_conditions=make_bc(times["ref"]),
                                    t0=t0.clone() if torch.is_tensor(t0) else t0, dt=dt)  # the reference's `total_t +=` writes into t0
        fake = FakeDiffusion(horizon)
        ro = MultiHorizonRollout(fake, horizon=horizon, window=window, num_predictions=members, autoregressive_steps=ar_steps,
                                 prediction_horizon=pred_h, prediction_inputs_noise=noise)
        torch.manual_seed(100 + trial)
        got = ro.evaluation_step(batch_d, "test", boundary_conditions=make_bc(times["mine"]), t0=t0, dt=dt, to_numpy=True)
        cfg = dict(horizon=horizon, window=window, members=members, batch=batch, ar_steps=ar_steps, pred_h=pred_h, noise=noise)
        assert list(got) == list(want), cfg
        for k in want:
            assert np.array_equal(got[k], want[k]), (cfg, k)
        assert fake.calls == exp.model.calls, cfg                      # same sampler calls: shapes and num_predictions
        assert len(times["ref"]) == len(times["mine"]) and all(
            (a == b) if isinstance(a, float) else torch.equal(a, b) for a, b in zip(times["ref"], times["mine"])), cfg
    assert seen_window == {1, 2, 3} and seen_partial >= 3
