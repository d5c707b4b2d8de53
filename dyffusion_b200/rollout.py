"""Stream-ordered drop-in for the reference's autoregressive evaluation loop (second half):
`AbstractMultiHorizonForecastingExperiment._evaluation_step` (`src/experiment_types/forecasting_multi_horizon.py:115-238`)
with `get_preds_at_t_for_batch` (:287-332), `get_inputs_and_extra_kwargs` / `get_extra_model_kwargs` (:344-388),
`transform_inputs` (:337-342), `BaseExperiment.predict` / `reshape_predictions` / `get_ensemble_inputs` /
`_reshape_ensemble_preds` (`src/experiment_types/_base_experiment.py:315-379, :503-567`) and the stacking done by
`test_step` (:240-262), for the DYffusion module.

What changes against the reference is WHERE things happen, not what is computed:
* one sampler invocation per autoregressive step (the reference caches it per step too, :296-313); its forecasts stay on
  the device, the boundary conditions are masked-write kernels on the sampler's own output buffer, and the hand-off to the
  next autoregressive step is a view of that buffer -- nothing is copied to the host between horizons (the reference
  converts every horizon with `torch_to_numpy`, :185-187) and nothing synchronises the stream;
* results are device tensors (`to_numpy=True` restores the reference's numpy dict);
* the caller's `batch["dynamics"]` is NOT multiplied by 1e6 after the first pass (:221 is a debugging guard; the targets
  come from a clone taken before the loop in the reference, so no result depends on it), and a tensor `t0` is not advanced in
  place (the reference's `total_t += ...` (:164) writes into the caller's `metadata["t"][:, 0]`).

The host logic is written against the `diffusion` object's `predict_forward` / `sample_loop` surface only, so the tests
drive it on CPU around the reference's own DYffusion module and compare with the reference loop bit for bit
(`tests/test_rollout_cpu.py`); on the GPU the object is `dyffusion_b200.diffusion.DYffusion` (native sampler)."""
from __future__ import annotations

import inspect
import math
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor


class MultiHorizonRollout:
    """Evaluation-time mirror of `MultiHorizonForecastingDYffusion` (forecasting_multi_horizon.py:391-424) around a
    DYffusion sampler.  Constructor arguments carry the reference's names: `horizon`/`window` are the datamodule's,
    `prediction_horizon` the datamodule's `prediction_horizon`, the rest `BaseExperiment` / experiment hparams."""

    CHANNEL_DIM = -3  # _base_experiment.py:59

    def __init__(self, diffusion, horizon: int, window: int = 1, num_predictions: int = 1,
                 autoregressive_steps: int = 0, prediction_horizon: Optional[int] = None,
                 prediction_timesteps: Optional[Sequence[float]] = None, prediction_inputs_noise: float = 0.0,
                 group=None):
        assert autoregressive_steps >= 0, f"Autoregressive steps must be >= 0, but is {autoregressive_steps}"
        if autoregressive_steps > 0:
            assert prediction_horizon is None, "Cannot use ``prediction_horizon`` with autoregressive_steps > 0"
        self.model = diffusion
        self.horizon, self.window = int(horizon), int(window)
        self.num_predictions = int(num_predictions)
        self.autoregressive_steps = int(autoregressive_steps)
        self._prediction_horizon = prediction_horizon
        self._prediction_timesteps = list(prediction_timesteps) if prediction_timesteps is not None else None
        self.inputs_noise = float(prediction_inputs_noise)
        self.stack_window_to_channel_dim = True
        self.group = group  # torch.distributed group: rows of every sampler call are sharded over its ranks
        timesteps = getattr(diffusion, "num_timesteps", None)
        hp = getattr(diffusion, "hparams", {})
        base_t = hp.get("timesteps", None) if hasattr(hp, "get") else None
        if base_t is not None:
            assert base_t == self.horizon, "diffusion timesteps must be equal to horizon"  # :395
        elif timesteps is not None:
            assert timesteps >= self.horizon
        if hasattr(diffusion, "interpolator"):  # :396-398
            diffusion.interpolator.hparams.num_predictions = self.num_predictions

    # ---------------------------------------------------------------- properties (:43-99)
    @property
    def horizon_range(self) -> List[int]:
        return list(np.arange(1, self.horizon + 1))

    @property
    def true_horizon(self) -> int:
        return self.horizon

    @property
    def prediction_timesteps(self) -> List[float]:
        return self._prediction_timesteps or self.horizon_range

    @prediction_timesteps.setter
    def prediction_timesteps(self, value: List[float]):
        assert max(value) <= self.horizon_range[-1], \
            f"Prediction range {value} exceeds horizon range {self.horizon_range}"
        self._prediction_timesteps = value

    @property
    def prediction_horizon(self) -> int:
        if self._prediction_horizon:
            return self._prediction_horizon
        return self.horizon * (self.autoregressive_steps + 1)

    @property
    def num_autoregressive_steps(self) -> int:
        n = self.autoregressive_steps
        if n == 0 and self.prediction_horizon is not None:
            n = max(1, math.ceil(self.prediction_horizon / self.true_horizon)) - 1
        return n

    def use_ensemble_predictions(self, split: str) -> bool:
        return self.num_predictions > 1 and split in ["val", "test", "predict"]

    # ---------------------------------------------------------------- inputs (:334-388, _base_experiment.py:503-538)
    def get_ensemble_inputs(self, inputs_raw, split: str, add_noise: bool = True, flatten_into_batch_dim: bool = True):
        """Member-major stacking `(N B) ...` of the reference (N outermost)."""
        if inputs_raw is None:
            return None
        if not self.use_ensemble_predictions(split):
            return inputs_raw
        n = self.num_predictions
        if isinstance(inputs_raw, dict):
            return {k: self.get_ensemble_inputs(v, split, add_noise, flatten_into_batch_dim) for k, v in inputs_raw.items()}
        if isinstance(inputs_raw, Sequence):
            return np.array([inputs_raw] * n)
        if add_noise:
            inputs = torch.stack([inputs_raw + self.inputs_noise * torch.randn_like(inputs_raw) for _ in range(n)], dim=0)
        else:
            inputs = inputs_raw.unsqueeze(0).expand(n, *inputs_raw.shape)  # the reference stacks n copies
        if flatten_into_batch_dim:
            inputs = inputs.reshape(n * inputs_raw.shape[0], *inputs_raw.shape[1:])
        return inputs

    def get_inputs_from_dynamics(self, dynamics: Tensor) -> Tensor:
        return dynamics[:, : self.window, ...]

    def transform_inputs(self, inputs: Tensor, split: str = None, ensemble: bool = True, **kwargs) -> Tensor:
        if self.stack_window_to_channel_dim and inputs.ndim == 5:
            b, w, c = inputs.shape[:3]
            inputs = inputs.reshape(b, w * c, *inputs.shape[3:])  # "b window c lat lon -> b (window c) lat lon"
        if ensemble:
            inputs = self.get_ensemble_inputs(inputs, split=split, **kwargs)
        return inputs

    def get_extra_model_kwargs(self, batch: Dict[str, Any], split: str, ensemble: bool,
                               is_autoregressive: bool = False) -> Dict[str, Any]:
        dshape = batch["dynamics"].shape
        extra: Dict[str, Any] = {}
        for k, v in batch.items():
            if k == "dynamics":
                continue
            if k == "metadata":
                extra[k] = v  # the reference stacks it N times (:359) for a callee that ignores it (_base_diffusion.py:48)
                continue
            no_channel = v.shape[1: self.CHANNEL_DIM] + v.shape[self.CHANNEL_DIM + 1:]
            time_varying = dshape[1: self.CHANNEL_DIM] + dshape[self.CHANNEL_DIM + 1:]
            if no_channel == time_varying:
                extra[k] = self.transform_inputs(self.get_inputs_from_dynamics(v), split=split, ensemble=ensemble,
                                                 add_noise=False)
            else:
                extra[k] = self.get_ensemble_inputs(v, split=split, add_noise=False) if ensemble else v
        return extra

    def get_inputs_and_extra_kwargs(self, batch, split: str = None, ensemble: bool = True,
                                    autoregressive_inputs: Optional[Tensor] = None):
        is_ar = autoregressive_inputs is not None
        if is_ar:
            inputs = autoregressive_inputs
        else:
            inputs = self.transform_inputs(self.get_inputs_from_dynamics(batch["dynamics"]), split=split, ensemble=True)
        return inputs, self.get_extra_model_kwargs(batch, split=split, ensemble=ensemble, is_autoregressive=is_ar)

    # ---------------------------------------------------------------- predict (_base_experiment.py:315-379, :540-567)
    def _sample(self, inputs: Tensor, **kwargs) -> Dict[str, Tensor]:
        if self.group is not None:
            from .distributed import sample_sharded
            kwargs.pop("metadata", None)
            static = kwargs.pop("condition", None)
            return sample_sharded(self.model, inputs.contiguous(), static_condition=static, group=self.group, **kwargs)
        return self.model.predict_forward(inputs, **kwargs)

    def predict(self, inputs: Tensor, num_predictions: Optional[int] = None, reshape_ensemble_dim: bool = True,
                **kwargs) -> Dict[str, Tensor]:
        n = num_predictions or self.num_predictions
        if hasattr(self.model, "sample_loop") and "num_predictions" in inspect.signature(self.model.sample_loop).parameters:
            kwargs["num_predictions"] = n
        for k, v in list(kwargs.items()):
            if torch.is_tensor(v) and not v.is_contiguous():
                kwargs[k] = v.contiguous()  # expanded ensemble views -> real rows for the engine
        results = self._sample(inputs if inputs.is_contiguous() else inputs.contiguous(), **kwargs)
        if torch.is_tensor(results):
            results = {"preds": results}
        # NB (:350-351): the reference restores hparams.num_predictions BEFORE reshaping, so the autoregressive calls
        # (num_predictions=1) are still un-stacked with the experiment's ensemble size
        return self.reshape_predictions(results, reshape_ensemble_dim)

    def reshape_predictions(self, results: Dict[str, Tensor], reshape_ensemble_dim: bool = True) -> Dict[str, Tensor]:
        n = self.num_predictions
        pred_keys = [k for k in results.keys() if "preds" in k]
        shape = results[pred_keys[0]].shape
        if reshape_ensemble_dim and shape[0] > 1:
            if n > 1 and shape[0] % n == 0:
                results = self._reshape_ensemble_preds(results, "predict")
                shape = results[pred_keys[0]].shape
            if 1 < n == shape[0] and len(shape) <= 4:
                for k in pred_keys:
                    results[k] = results[k].unsqueeze(1)
        return results

    def _reshape_ensemble_preds(self, results: Dict[str, Tensor], split: str) -> Dict[str, Tensor]:
        n = self.num_predictions
        if self.use_ensemble_predictions(split):
            for key in results:
                if "targets" not in key and "true" not in key:
                    b = results[key].shape[0]
                    assert b % n == 0, \
                        f"key={key}: b % #ens_mems = {b} % {n} != 0 ...Did you forget to create the input ensemble?"
                    results[key] = results[key].reshape(n, max(1, b // n), *results[key].shape[1:])
        return results

    # ---------------------------------------------------------------- one horizon of one AR step (:287-332)
    def get_preds_at_t_for_batch(self, batch, horizon, split: str, autoregressive_inputs: Optional[Tensor] = None,
                                 ensemble: bool = False, **kwargs) -> Dict[str, Tensor]:
        assert 0 < horizon <= self.true_horizon, f"horizon={horizon} must be in [1, {self.true_horizon}]"
        if horizon == self.prediction_timesteps[0]:
            if self.prediction_timesteps != self.horizon_range:
                self.model.hparams.prediction_timesteps = [p_h for p_h in self.prediction_timesteps]
            inputs, extra = self.get_inputs_and_extra_kwargs(batch, split=split, ensemble=ensemble,
                                                            autoregressive_inputs=autoregressive_inputs)
            with torch.no_grad():
                self._current_preds = self.predict(inputs, **extra, **kwargs)
        preds_key = f"t{horizon}_preds"
        results = {k: self._current_preds.pop(k) for k in list(self._current_preds.keys()) if preds_key in k}
        if horizon == self.horizon_range[-1]:
            assert all(["preds" not in k for k in self._current_preds.keys()]), (
                f'preds_key={preds_key} must be the only key containing "preds" in last prediction. '
                f"Got: {list(self._current_preds.keys())}")
            results = {**results, **self._current_preds}
            del self._current_preds
        return results

    # ---------------------------------------------------------------- the loop (:115-238)
    @torch.no_grad()
    def evaluation_step(self, batch: Dict[str, Any], split: str = "test", return_outputs=True,
                        boundary_conditions: Callable = None, t0=0.0, dt=1.0, autoregressive: bool = True,
                        to_numpy: bool = False) -> Dict[str, Any]:
        """Returns the reference's `return_dict`: `t{k}_preds` ((N, B, C, H, W), or (B, C, H, W) without an ensemble)
        and `t{k}_targets` for every total horizon k -- device tensors unless `to_numpy`.  `autoregressive=False` is the
        reference's first validation loader (one pass, :135-136)."""
        return_dict: Dict[str, Any] = dict()
        conv = (lambda x: None if x is None else x.detach().cpu().numpy()) if to_numpy else (lambda x: x)
        dynamics = batch["dynamics"]  # the reference clones; nothing below writes to it
        if not autoregressive:
            n_outer_loops = 1
        else:
            assert split in ["val", "test", "predict"]
            n_outer_loops = self.num_autoregressive_steps + 1
            if dynamics.shape[1] < self.prediction_horizon:
                raise ValueError(f"Prediction horizon {self.prediction_horizon} is larger than {dynamics.shape}[1]")
        autoregressive_inputs = None
        total_t = t0
        predicted_range_last = [0.0] + self.prediction_timesteps[:-1]
        ar_window_steps_t = self.horizon_range[-self.window:]
        for ar_step in range(n_outer_loops):
            ar_window_steps = []
            for t_step_last, t_step in zip(predicted_range_last, self.prediction_timesteps):
                total_horizon = ar_step * self.true_horizon + t_step
                if total_horizon > self.prediction_horizon:
                    break
                pr_kwargs = {} if autoregressive_inputs is None else {"num_predictions": 1}
                results = self.get_preds_at_t_for_batch(batch, t_step, split, autoregressive_inputs, ensemble=True,
                                                        **pr_kwargs)
                total_t = total_t + dt * (t_step - t_step_last)
                if float(total_horizon).is_integer():
                    targets = dynamics[:, self.window + int(total_horizon) - 1, ...]
                else:
                    targets = None
                pk = f"t{t_step}_preds"
                if boundary_conditions is not None:
                    results[pk] = boundary_conditions(preds=results[pk], targets=targets,
                                                      metadata=batch.get("metadata", None), time=total_t)
                preds = results.pop(pk)
                if return_outputs in [True, "all"]:
                    return_dict[f"t{total_horizon}_targets"] = conv(targets)
                    return_dict[f"t{total_horizon}_preds"] = conv(preds)
                if return_outputs == "all":
                    return_dict.update({k.replace(f"t{t_step}", f"t{total_horizon}"): conv(v)
                                        for k, v in results.items()})
                if t_step in ar_window_steps_t:
                    ar_window_steps += [preds.reshape(-1, *preds.shape[-3:]).unsqueeze(1)]
            if ar_step < n_outer_loops - 1:
                if len(ar_window_steps) == 1:
                    autoregressive_inputs = ar_window_steps[0]  # a view of the sampler's output buffer (window == 1)
                else:
                    autoregressive_inputs = torch.cat(ar_window_steps, dim=1)
                autoregressive_inputs = self.transform_inputs(autoregressive_inputs, split=split, ensemble=False)
        return return_dict

    # ---------------------------------------------------------------- test_step's stacking (:240-262)
    def stack_trajectory(self, results: Dict[str, Any]):
        """`(predicted_trajectory, true_trajectory)` as `test_step` builds them (:246-249): horizons stacked on axis -5,
        i.e. (N, T, B, C, H, W) and (T, B, C, H, W) -- the operands of `evaluate_ensemble_prediction`."""
        ts = range(1, self.prediction_horizon + 1)
        stack = np.stack if isinstance(results[f"t{ts[0]}_preds"], np.ndarray) else torch.stack
        return (stack([results[f"t{t}_preds"] for t in ts], -5), stack([results[f"t{t}_targets"] for t in ts], -5))

    def test_step(self, batch: Dict[str, Any], boundary_conditions: Callable = None, t0=0.0, dt=1.0) -> Dict[str, Any]:
        """The physical-systems branch of the reference's `test_step` (:240-262) with everything on the device: rollout,
        boundary conditions, stacking and the ensemble metrics (`dyffusion_b200.metrics`); only the per-timestep metric
        table comes back to the host.  Returns `{"ssr", "crps", "mse"}` arrays of length `prediction_horizon`."""
        from .metrics import evaluate_ensemble_prediction
        results = self.evaluation_step(batch, "test", True, boundary_conditions, t0, dt)
        preds, targets = self.stack_trajectory(results)
        return evaluate_ensemble_prediction(preds, targets, mean_over_samples=False)
