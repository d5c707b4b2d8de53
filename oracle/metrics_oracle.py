"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64) of the reference's ensemble evaluation
(`src/utilities/evaluation.py`), the checker for `dyf_ensemble_metrics`.  Only tests/ may import this.

* `evaluate_ensemble_prediction` follows evaluation.py:10-80 (ensemble-mean MSE :40-44, per-member MSE :47-53, CRPS :66,
  spread-skill ratio :69-71) and `evaluate_ensemble_spread_skill_ratio` :100-120 (population variance over members,
  averaged, square-rooted, divided by the RMSE of the ensemble mean).
* `crps_ensemble` restates the algorithm the reference reaches through `xskillscore.crps_ensemble` (evaluation.py:83-97)
  -> `properscoring.crps_ensemble` (`properscoring/_crps.py::_crps_ensemble_vectorized`, properscoring 0.1; xskillscore is
  un-pinned in the reference's setup.py:96): members sorted, then the integral of (F_ens - F_obs)^2 accumulated member by
  member with equal weights.

Pinning: MSE, per-member MSE, spread-skill ratio and all shape / `mean_over_samples` bookkeeping are pinned against the
reference's OWN `src/utilities/evaluation.py`, imported unmodified with stub `xarray` / `xskillscore` modules
(tests/test_metrics_cpu.py::test_oracle_equals_the_reference_evaluation_module).  The CRPS kernel itself lives in third-party
code that is absent here (xskillscore, un-pinned in the reference's setup.py:96 -> properscoring 0.1
`_crps_ensemble_vectorized`): for that one function parity stays "unpinned against the library"; it is pinned to the published
algorithm restated below, to the closed form it integrates, CRPS = E|X - y| - 1/2 E|X - X'| (brute force, float64), and to
hand-computed cases.
"""
from __future__ import annotations

import numpy as np


def crps_ensemble(observations: np.ndarray, forecasts: np.ndarray) -> np.ndarray:
    """observations: [...]; forecasts: [..., members] -> CRPS [...] (float64).  Literal restatement of properscoring's
    `_crps_ensemble_vectorized` with equal member weights."""
    obs = np.asarray(observations, dtype=np.float64)
    fc = np.sort(np.asarray(forecasts, dtype=np.float64), axis=-1)
    n = fc.shape[-1]
    weight = 1.0 / n
    out = np.empty(obs.shape, dtype=np.float64)
    flat_obs, flat_fc, flat_out = obs.reshape(-1), fc.reshape(-1, n), out.reshape(-1)
    for idx in range(flat_obs.shape[0]):
        observation = flat_obs[idx]
        obs_cdf = 0.0
        forecast_cdf = 0.0
        prev_forecast = 0.0
        integral = 0.0
        forecast = 0.0
        for k in range(n):
            forecast = flat_fc[idx, k]
            if obs_cdf == 0 and observation < forecast:
                integral += (observation - prev_forecast) * forecast_cdf ** 2
                integral += (forecast - observation) * (forecast_cdf - 1) ** 2
                obs_cdf = 1.0
            else:
                integral += (forecast - prev_forecast) * (forecast_cdf - obs_cdf) ** 2
            forecast_cdf += weight
            prev_forecast = forecast
        if obs_cdf == 0:
            integral += observation - forecast
        flat_out[idx] = integral
    return out


def crps_energy_form(observations: np.ndarray, forecasts: np.ndarray) -> np.ndarray:
    """CRPS = E|X - y| - 1/2 E|X - X'| by brute force (the identity the sorted algorithm integrates)."""
    obs = np.asarray(observations, dtype=np.float64)[..., None]
    fc = np.asarray(forecasts, dtype=np.float64)
    return np.abs(fc - obs).mean(-1) - 0.5 * np.abs(fc[..., :, None] - fc[..., None, :]).mean((-1, -2))


def evaluate_ensemble_prediction(predictions, targets, also_per_member_metrics=False, mean_over_samples=True):
    """predictions [members, samples, *], targets [samples, *] -> {"ssr", "crps", "mse"[, "mse_per_mem", "mse_per_mem_mean"]}
    (evaluation.py:10-80)."""
    predictions = np.asarray(predictions, dtype=np.float64)
    targets = np.asarray(targets, dtype=np.float64)
    assert predictions.shape[1] == targets.shape[0]
    if predictions.ndim == 3:
        predictions = predictions[:, :, None]
    if targets.ndim == 2:
        targets = targets[:, None]
    mean_preds = predictions.mean(axis=0)
    mean_dims = tuple(range(mean_preds.ndim)) if mean_over_samples else tuple(range(1, mean_preds.ndim))
    mse = np.mean((mean_preds - targets) ** 2, axis=mean_dims)
    out = {"mse": mse}
    if also_per_member_metrics:
        diff = predictions - targets
        out["mse_per_mem"] = np.mean(diff ** 2, axis=tuple(range(1, predictions.ndim)))
        out["mse_per_mem_mean"] = np.mean(out["mse_per_mem"])
    crps = crps_ensemble(targets, np.moveaxis(predictions, 0, -1)).mean(axis=mean_dims)
    variance = np.var(predictions, axis=0).mean(axis=mean_dims)
    out["ssr"] = np.sqrt(variance) / np.sqrt(mse)
    out["crps"] = float(crps) if mean_over_samples else crps
    return out
