"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's autoregressive evaluation loop for the DYffusion module
(`AbstractMultiHorizonForecastingExperiment._evaluation_step`, `src/experiment_types/forecasting_multi_horizon.py:115-238`,
with the helpers it calls: `get_preds_at_t_for_batch` :287-332, `get_inputs_and_extra_kwargs` :372-388,
`BaseExperiment.predict` / `reshape_predictions` `_base_experiment.py:315-379`, `get_ensemble_inputs` :503-538), the checker
for `dyffusion_b200.rollout` (second half).  Only tests/ may import this.

Written as ONE flat function over a `sample(initial_condition [R, window*C, H, W], static [R, Cs, H, W] | None) -> dict`
callable (window stacked in channels, rows member-major), numpy results like the reference's `return_dict`.

Pinned: `tests/test_rollout_cpu.py` runs the reference's own `_evaluation_step` (imported through oracle/ref_shims.py in the
build container) with the reference's DYffusion module and boundary-condition method and compares bit for bit; the same run
is committed as `tests/golden/rollout_*.pt` by `tests/golden/make_rollout_golden.py` for boxes without /root/reference."""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch


def evaluation_step(sample: Callable, batch: Dict, *, horizon: int, window: int = 1, num_predictions: int = 1,
                    autoregressive_steps: int = 0, prediction_horizon: Optional[int] = None,
                    prediction_timesteps: Optional[Sequence[float]] = None, boundary_conditions: Callable = None,
                    t0=0.0, dt=1.0) -> Dict[str, np.ndarray]:
    N = num_predictions
    dynamics = batch["dynamics"].clone()                                   # :132
    B = dynamics.shape[0]
    horizon_range = list(np.arange(1, horizon + 1))                        # :43-45
    steps = list(prediction_timesteps) if prediction_timesteps else horizon_range  # :55-58
    pred_h = prediction_horizon or horizon * (autoregressive_steps + 1)    # :95-99
    n_ar = autoregressive_steps                                            # :70-75
    if n_ar == 0 and pred_h is not None:
        n_ar = max(1, math.ceil(pred_h / horizon)) - 1
    if dynamics.shape[1] < pred_h:                                         # :139-140
        raise ValueError(f"Prediction horizon {pred_h} is larger than {dynamics.shape}[1]")

    static = batch.get("condition", None)
    if static is not None and N > 1:                                       # :369-370 -> get_ensemble_inputs (N B)
        static = torch.stack([static for _ in range(N)], dim=0).flatten(0, 1)

    out: Dict[str, np.ndarray] = {}
    ar_inputs = None
    total_t = t0
    last_steps = [0.0] + steps[:-1]
    ar_window_t = horizon_range[-window:]
    for ar in range(n_ar + 1):
        window_preds, current = [], None
        for t_last, t_step in zip(last_steps, steps):
            total_h = ar * horizon + t_step
            if total_h > pred_h:
                break
            if t_step == steps[0]:                                         # one sampler call per AR step (:296-313)
                if ar_inputs is None:
                    x = dynamics[:, :window].flatten(1, 2)                 # :334-339
                    if N > 1:
                        # :523-535 with prediction_inputs_noise = 0: the N draws are still taken from the torch RNG
                        x = torch.stack([x + 0.0 * torch.randn_like(x) for _ in range(N)], dim=0).flatten(0, 1)
                else:
                    x = ar_inputs
                with torch.no_grad():
                    current = dict(sample(x, static))
                for k in list(current):                                    # reshape_predictions with the BASE ensemble size
                    v = current[k]
                    if v.shape[0] > 1 and N > 1 and v.shape[0] % N == 0:
                        current[k] = v.reshape(N, max(1, v.shape[0] // N), *v.shape[1:])
            preds = current.pop(f"t{t_step}_preds")
            total_t = total_t + dt * (t_step - t_last)                     # :164
            targets = dynamics[:, window + int(total_h) - 1] if float(total_h).is_integer() else None
            if boundary_conditions is not None:                            # :175-182
                preds = boundary_conditions(preds=preds, targets=targets, metadata=batch.get("metadata", None), time=total_t)
            out[f"t{total_h}_targets"] = None if targets is None else targets.numpy()
            out[f"t{total_h}_preds"] = preds.detach().numpy()
            if t_step in ar_window_t:                                      # :195-198
                window_preds.append(preds.reshape(-1, *preds.shape[-3:]).unsqueeze(1))
        if ar < n_ar:                                                      # :217-220
            ar_inputs = torch.cat(window_preds, dim=1).flatten(1, 2)
    return out
