"""On-device drop-in for the reference's ensemble evaluation (`src/utilities/evaluation.py:10-120`).

`evaluate_ensemble_prediction` keeps the reference's name, arguments and result keys ("ssr", "crps", "mse"
[, "mse_per_mem", "mse_per_mem_mean"]) but takes the CUDA tensors the sampler produced instead of numpy copies of every
horizon (the reference converts with `torch_to_numpy` first, `forecasting_multi_horizon.py:185-187`): the arithmetic runs
in `dyf_ensemble_metrics` and only a [samples, 3] table of sums comes back to the host.  No CPU / PyTorch fallback."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import engine as E


def evaluate_ensemble_prediction(predictions: torch.Tensor, targets: torch.Tensor, ensemble_dim: int = 0,
                                 also_per_member_metrics: bool = False, mean_over_samples: bool = True) -> Dict[str, object]:
    """predictions: (n_members, n_samples, *), targets: (n_samples, *) -- CUDA tensors (evaluation.py:10-80)."""
    if ensemble_dim != 0:
        raise ValueError("ensemble_dim must be 0 (the reference indexes predictions.shape[1] as the sample axis)")
    assert predictions.shape[1] == targets.shape[ensemble_dim], \
        f"predictions.shape[1] ({predictions.shape[1]}) != targets.shape[0] ({targets.shape[ensemble_dim]})"
    if not (predictions.is_cuda and targets.is_cuda):
        raise E.EngineError("dyffusion_b200.metrics has no CPU path: pass the CUDA tensors the sampler returned")
    n, s = predictions.shape[:2]
    preds = predictions.reshape(n, s, -1).float()
    tgts = targets.reshape(s, -1).float()
    inner = preds.shape[2]
    per_sample, member = E.ensemble_metrics(preds, tgts, per_member=also_per_member_metrics)
    sums = per_sample.cpu().numpy()  # [samples, 3]: crps, squared error of the ensemble mean, member variance (sums over inner)
    if mean_over_samples:
        crps, mse, var = (sums.sum(axis=0) / (s * inner)).tolist()
        out: Dict[str, object] = {"ssr": np.sqrt(var) / np.sqrt(mse), "crps": float(crps), "mse": np.float64(mse)}
    else:
        per = sums / inner
        out = {"ssr": np.sqrt(per[:, 2]) / np.sqrt(per[:, 1]), "crps": per[:, 0], "mse": per[:, 1]}
    if also_per_member_metrics:
        out["mse_per_mem"] = member.cpu().numpy()
        out["mse_per_mem_mean"] = np.mean(out["mse_per_mem"])
    return out
