"""Evaluation-side mirror of the reference's `InterpolationExperiment` (`src/experiment_types/interpolation.py:12-167`) around an
engine backbone: the caller on the INTERPOLATOR side of the hot path (stage 1 of DYffusion; BASELINE.json configs[0] is one
forward of this experiment).  Same input construction (`get_inputs_from_dynamics` :132-146: the window stacked in channels +
the last frame), same ensemble stacking, same per-time loop and result keys (`_evaluation_step` :68-130), same loss batch
(`get_loss` :148-167: one random intermediate time per row); results stay on the device.

Host logic only -- the arithmetic is the backbone's `predict_forward` / `get_loss` -- and pinned on CPU by driving the
reference's own interpolator module through it (`tests/test_interpolation_cpu.py`)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np
import torch
from torch import Tensor


class InterpolationEvaluation:
    def __init__(self, model, horizon: int, window: int = 1, num_predictions: int = 1, stack_window_to_channel_dim: bool = True,
                 enable_inference_dropout: bool = True, prediction_inputs_noise: float = 0.0):
        assert horizon >= 2, "horizon must be >=2 for interpolation experiments"
        self.model, self.horizon, self.window = model, int(horizon), int(window)
        self.num_predictions = int(num_predictions)
        self.stack_window_to_channel_dim = stack_window_to_channel_dim
        self.enable_inference_dropout = enable_inference_dropout  # module/_base_experiment_config.yaml; interpolation runs: True
        self.inputs_noise = float(prediction_inputs_noise)

    @property
    def horizon_range(self) -> List[int]:
        return list(np.arange(1, self.horizon))  # interpolate between t = 0 and t = horizon (:22-27)

    def use_ensemble_predictions(self, split: str) -> bool:
        return self.num_predictions > 1 and split in ["val", "test", "predict"]

    def get_ensemble_inputs(self, inputs_raw: Optional[Tensor], split: str, add_noise: bool = True):
        if inputs_raw is None or not self.use_ensemble_predictions(split):
            return inputs_raw
        n = self.num_predictions
        if isinstance(inputs_raw, dict):
            return {k: self.get_ensemble_inputs(v, split, add_noise) for k, v in inputs_raw.items()}
        if add_noise:
            stacked = torch.stack([inputs_raw + self.inputs_noise * torch.randn_like(inputs_raw) for _ in range(n)], dim=0)
        else:
            stacked = torch.stack([inputs_raw for _ in range(n)], dim=0)
        return stacked.reshape(n * inputs_raw.shape[0], *inputs_raw.shape[1:])  # "N B ... -> (N B) ..."

    def get_inputs_from_dynamics(self, dynamics: Tensor) -> Tensor:
        assert dynamics.shape[1] == self.window + self.horizon, "dynamics must have shape (b, t, c, h, w)"
        past, last = dynamics[:, : self.window, ...], dynamics[:, -1, ...]
        if self.stack_window_to_channel_dim:
            past = past.reshape(past.shape[0], past.shape[1] * past.shape[2], *past.shape[3:])
        else:
            last = last.unsqueeze(1)
        return torch.cat([past, last], dim=1)

    def predict(self, inputs: Tensor, **kwargs) -> Dict[str, Tensor]:
        """`BaseExperiment.predict` (`_base_experiment.py:315-356`): forward, then the ensemble un-stacking `(N B) ... -> N B ...`."""
        with torch.no_grad():
            preds = self.model.predict_forward(inputs, **kwargs)
        n = self.num_predictions
        if preds.shape[0] > 1 and n > 1 and preds.shape[0] % n == 0:
            preds = preds.reshape(n, max(1, preds.shape[0] // n), *preds.shape[1:])
        return {"preds": preds}

    @torch.no_grad()
    def evaluation_step(self, batch: Dict[str, Any], split: str = "val", return_only_preds_and_targets: bool = False) -> Dict[str, Tensor]:
        """`evaluation_step` (`_base_experiment.py:484-493`: the whole step runs under the inference-dropout scope) around
        `_evaluation_step` (interpolation.py:68-130)."""
        dynamics = batch["dynamics"]
        inputs = self.get_ensemble_inputs(self.get_inputs_from_dynamics(dynamics), split)
        extra = {k: self.get_ensemble_inputs(v, split, add_noise=False) for k, v in batch.items() if k != "dynamics"}
        out: Dict[str, Tensor] = {}
        with self.model.inference_dropout_scope(condition=bool(self.enable_inference_dropout)):
            for t_step in self.horizon_range:
                targets = dynamics[:, self.window + t_step - 1, ...]
                time = torch.full((inputs.shape[0],), t_step, device=inputs.device, dtype=torch.long)
                results = self.predict(inputs, time=time, **extra)
                out[f"t{t_step}_preds"] = results["preds"]
                out[f"t{t_step}_targets"] = targets
        return out

    def get_loss(self, batch: Dict[str, Any]) -> Tensor:
        """One random intermediate time per row (:148-167); the criterion is the backbone's."""
        dynamics = batch["dynamics"]
        inputs = self.get_inputs_from_dynamics(dynamics)
        b, dev = dynamics.shape[0], dynamics.device
        possible = torch.tensor(self.horizon_range, device=dev, dtype=torch.long)
        t = possible[torch.randint(len(possible), (b,), device=dev, dtype=torch.long)]
        targets = dynamics[torch.arange(b), self.window + t - 1, ...]
        return self.model.get_loss(inputs=inputs, targets=targets, time=t, **{k: v for k, v in batch.items() if k != "dynamics"})
