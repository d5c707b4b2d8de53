"""Drop-in for `src.diffusion.dyffusion.DYffusion` (reference src/diffusion/dyffusion.py:17-567 on top of
src/diffusion/_base_diffusion.py:13-117): same constructor arguments, attributes and `sample / sample_loop /
predict_forward` contract.  The sampling loop itself (`sample_loop`, :335-426) runs natively in the CUDA engine
(`dyf_sampler_run`): forecaster and interpolator launches are enqueued back to back on the current stream with no
Python in between, the two interpolator evaluations of a cold-sampling step and the refinement calls are batched.

A Python-driven loop over the same engine-backed networks is kept for the options the native loop does not cover
(`log_every_t`, schedules that stop early, foreign interpolator objects); it launches the same CUDA kernels.
"""
from __future__ import annotations

import inspect
import math
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch
from torch import Tensor, nn

from .. import engine as E
from .._base import BaseModel, EngineBackbone
from .schedule import DiffusionSchedule


def _freeze(module: nn.Module) -> nn.Module:
    """src/utilities/utils.py:553-557."""
    for p in module.parameters():
        p.requires_grad = False
    return module.eval()


class InterpolatorHandle(nn.Module):
    """Minimal stand-in for the reference's `InterpolationExperiment` (src/experiment_types/interpolation.py:12-167)
    around an engine backbone: exposes exactly what `DYffusion` reads -- `.model`, `.window`, `.true_horizon`,
    `.hparams.num_predictions`, `.predict(...)`, `.inference_dropout_scope(...)`.  This is the offline route the
    reference lacks: build the interpolator backbone, `load_state_dict`, wrap it here."""

    def __init__(self, model: EngineBackbone, horizon: int, window: int = 1):
        super().__init__()
        self.model = model
        self.horizon, self.window = int(horizon), int(window)
        self.hparams = type("HP", (dict,), {"__getattr__": dict.__getitem__, "__setattr__": dict.__setitem__})(
            num_predictions=1)

    @property
    def true_horizon(self) -> int:
        return self.horizon

    def inference_dropout_scope(self, condition: bool = None, context=None):
        return self.model.inference_dropout_scope(bool(condition))

    def predict(self, inputs: Tensor, num_predictions: Optional[int] = None, reshape_ensemble_dim: bool = True,
                **kwargs) -> Dict[str, Tensor]:
        return {"preds": self.model.predict_forward(inputs, **kwargs)}


class DYffusion(BaseModel):
    def __init__(self, model: BaseModel = None, timesteps: int = None,
                 interpolator: Optional[nn.Module] = None, interpolator_run_id: Optional[str] = None,
                 interpolator_local_checkpoint_path: Optional[str] = None,
                 interpolator_wandb_ckpt_filename: Optional[str] = None,
                 lambda_reconstruction: float = 1.0, lambda_reconstruction2: float = 0.0,
                 forward_conditioning: str = "data", schedule: str = "before_t1_only",
                 additional_interpolation_steps: int = 0, additional_interpolation_steps_factor: int = 0,
                 interpolate_before_t1: bool = False, sampling_type: str = "cold",
                 sampling_schedule: Union[List[float], str] = None, time_encoding: str = "dynamics",
                 refine_intermediate_predictions: bool = False,
                 prediction_timesteps: Optional[Sequence[float]] = None,
                 enable_interpolator_dropout: Union[bool, str] = True,
                 use_cold_sampling_for_last_step: bool = False, log_every_t: Union[str, int] = None,
                 sampling_timesteps: int = None, max_rows_per_call: int = 0, cuda_graph: Optional[bool] = None,
                 **kwargs):
        super().__init__(**kwargs)
        if model is None:
            raise ValueError("Arg ``model`` is missing... Please provide a backbone model for the diffusion model "
                             "(e.g. a Unet)")
        sampling_schedule = None if sampling_schedule == "None" else sampling_schedule
        loc = dict(locals())
        for k in ("self", "model", "interpolator", "kwargs", "__class__"):
            loc.pop(k, None)
        self._record_hparams(loc)
        self.model = model
        self.spatial_shape = model.spatial_shape
        self.num_input_channels = model.num_input_channels
        self.num_output_channels = model.num_output_channels
        self.num_conditional_channels = model.num_conditional_channels

        if forward_conditioning not in ("data", "none", "data+noise"):
            raise ValueError(f"Invalid value for forward_conditioning: {forward_conditioning}")
        if enable_interpolator_dropout not in (True, False):
            raise ValueError(f"Invalid value for enable_interpolator_dropout: {enable_interpolator_dropout}")
        self._sched = DiffusionSchedule(timesteps, schedule, additional_interpolation_steps,
                                        additional_interpolation_steps_factor, interpolate_before_t1)
        self.num_timesteps = self._sched.num_timesteps
        self.additional_diffusion_steps = self._sched.additional_diffusion_steps
        self.dynamical_steps = self._sched.dynamical_steps
        self.artificial_interpolation_steps = self._sched.artificial_interpolation_steps
        self.i_to_diffusion_step = self._sched.i_to_diffusion_step
        self.enable_interpolator_dropout = enable_interpolator_dropout
        self.full_sampling_schedule = list(range(self.num_timesteps))
        self.sampling_schedule = sampling_schedule or self.full_sampling_schedule

        # ---- the interpolator (reference :461-478; plus the offline routes of SURVEY.md F10)
        if interpolator is None:
            if interpolator_run_id is not None:
                raise NotImplementedError("interpolator_run_id needs wandb/network access; pass `interpolator=` "
                                          "(a module) or wrap a loaded backbone in InterpolatorHandle")
            raise ValueError("Provide either model_checkpoint, model_checkpoint_path or wandb_run_id")
        if isinstance(interpolator, EngineBackbone):
            interpolator = InterpolatorHandle(interpolator, horizon=timesteps)
        if interpolator_local_checkpoint_path is not None:
            from ..checkpoint import split_state_dict  # Lightning .ckpt of the interpolator's own run, or a bare state dict
            state = torch.load(interpolator_local_checkpoint_path, map_location="cpu", weights_only=False)
            state = state.get("state_dict", state)
            if any(k.startswith("model.") for k in state):
                state = split_state_dict(state)["model"]
            interpolator.model.load_state_dict(state)
        self.interpolator = _freeze(interpolator)
        self.interpolator_window = self.interpolator.window
        self.interpolator_horizon = self.interpolator.true_horizon
        last = self.diffusion_step_to_interpolation_step(self.num_timesteps - 1)
        if self.interpolator_horizon != last + 1:
            raise ValueError(f"interpolator horizon {self.interpolator_horizon} must be equal to the "
                             f"last interpolation step+1=i_N=i_{self.num_timesteps - 1}={last + 1}")
        self._native_cache: Dict[Any, E.SamplerHandle] = {}
        self._calls = 0

    # ------------------------------------------------------------------ schedule surface (reference :97-138, :241-333)
    @property
    def diffusion_steps(self) -> List[int]:
        return list(range(self.num_timesteps))

    def diffusion_step_to_interpolation_step(self, diffusion_step):
        if torch.is_tensor(diffusion_step):
            d = diffusion_step
            assert (0 <= d).all() and (d <= self.num_timesteps - 1).all(), f"diffusion_step out of range: {d}"
            if self._sched.kind == "linear":
                return (d + self._sched.offset) / (self._sched.factor + 1)
            k = self._sched.aux_steps
            return torch.where(d >= k + 1, (d - k).float(), d / (k + 1))
        return self._sched.interpolation_time(diffusion_step)

    @property
    def sampling_schedule(self) -> List[Union[int, float]]:
        return self._sampling_schedule

    @sampling_schedule.setter
    def sampling_schedule(self, schedule):
        self._sampling_schedule = self._sched.parse_sampling_schedule(schedule, warn=self.log_text.warning)

    # ------------------------------------------------------------------ reference routing (_base_diffusion.py:48-68)
    def predict_forward(self, inputs, condition=None, metadata: Any = None, **kwargs):
        if inputs is not None and condition is not None:
            kwargs["static_condition"] = condition
        return self.sample(inputs, **kwargs)

    @torch.no_grad()
    def sample(self, initial_condition, num_samples=1, **kwargs):
        return self.sample_loop(initial_condition, **kwargs)[1]

    def p_losses(self, xt_last: Tensor, condition: Tensor, t: Tensor, static_condition: Tensor = None):
        """The DYffusion objective (reference :496-567): forecaster loss on x_t (the initial condition for t = 0, an
        interpolator sample for t > 0) plus, weighted `lambda_reconstruction2`, the loss of a second forecast from the
        interpolation between the initial condition and the FIRST forecast at t + 1 (rows with t <= N - 2).

        Forward only on the CUDA engine: this is the reference's validation loss (`val/loss`, `val/loss_forward[2]`, computed
        under `torch.no_grad()` in eval mode).  The engine has no backward kernels yet, so calling it
        with gradients enabled in train mode raises from the backbone's `forward`; around torch backbones (foreign
        modules) it is differentiable as written."""
        lam1, lam2 = self.hparams.lambda_reconstruction, self.hparams.lambda_reconstruction2
        sub = lambda v, m: None if v is None else v[m]
        x_t = condition.clone()
        later = t > 0
        if later.any():  # rows with t > 0 start from an interpolator sample at step t
            x_ipol = self.q_sample(x_end=condition[later], x0=xt_last[later], t=t[later],
                                   static_condition=sub(static_condition, later), num_predictions=1)
            x_t[later] = x_ipol.to(x_t.dtype)
        pred = self.predict_x_last(condition=condition, x_t=x_t, t=t, static_condition=static_condition)
        loss_forward = self.criterion(pred, xt_last)
        has_next = t <= self.num_timesteps - 2
        loss_forward2 = 0.0
        if lam2 > 0 and has_next.any():
            t2 = t[has_next] + 1
            cond2, static2 = condition[has_next], sub(static_condition, has_next)
            x_ipol2 = self.q_sample(x_end=cond2, x0=pred[has_next], t=t2, static_condition=static2, num_predictions=1)
            pred2 = self.predict_x_last(condition=cond2, x_t=x_ipol2, t=t2, static_condition=static2)
            loss_forward2 = self.criterion(pred2, xt_last[has_next])
        prefix = "train" if self.training else "val"
        return {"loss": lam1 * loss_forward + lam2 * loss_forward2, f"{prefix}/loss_forward": loss_forward,
                f"{prefix}/loss_forward2": loss_forward2}

    def forward(self, inputs, targets=None, condition=None, time=None):
        """reference _base_diffusion.py:81-106: a random diffusion step per row unless `time` is given."""
        b = (targets if targets is not None else inputs).shape[0]
        t = time if time is not None else torch.randint(0, self.num_timesteps, (b,), device=inputs.device,
                                                         dtype=torch.long)
        return self.p_losses(targets, condition=inputs, t=t, static_condition=condition)

    def get_loss(self, inputs, targets, metadata: Any = None, **kwargs):
        return self(inputs, targets, **kwargs)

    # ------------------------------------------------------------------ single network calls (reference :140-239, :480-494)
    def _forecaster_time(self, s):
        enc = self.hparams.time_encoding
        if enc == "discrete":
            return s
        if enc == "normalized":
            return s / self.num_timesteps
        if enc == "dynamics":
            return self.diffusion_step_to_interpolation_step(s)
        raise ValueError(f"Invalid time_encoding: {enc}")

    def predict_x_last(self, condition: Tensor, x_t: Tensor, t: Tensor, is_sampling: bool = False,
                       static_condition: Optional[Tensor] = None):
        assert (0 <= t).all() and (t <= self.num_timesteps - 1).all(), f"Invalid timestep: {t}"
        kind = self.hparams.forward_conditioning
        if kind == "data":
            cond = condition
        elif kind == "none":
            cond = None
        else:
            w = (t / (self.num_timesteps - 1)).view(condition.shape[0], *[1] * (condition.ndim - 1))
            cond = w * condition + (1 - w) * torch.randn_like(condition)
        if static_condition is not None:
            cond = static_condition if cond is None else torch.cat([cond, static_condition], dim=1)
        return self.model.predict_forward(x_t, time=self._forecaster_time(t), condition=cond)

    def q_sample(self, x0, x_end, t: Optional[Tensor], interpolation_time: Optional[Tensor] = None,
                 is_artificial_step: bool = True, **kwargs) -> Tensor:
        assert t is None or interpolation_time is None, "Either t or interpolation_time must be None."
        t = interpolation_time if t is None else self.diffusion_step_to_interpolation_step(t)
        enable = bool(self.training or self.enable_interpolator_dropout)
        with self.interpolator.inference_dropout_scope(condition=enable):
            return self._interpolate(initial_condition=x_end, x_last=x0, t=t, **kwargs)

    def _interpolate(self, initial_condition: Tensor, x_last: Tensor, t: Tensor,
                     static_condition: Optional[Tensor] = None, **kwargs):
        assert (0 < t).all() and (t < self.interpolator_horizon).all(), \
            f"interpolate time must be in (0, {self.interpolator_horizon}), got {t}"
        kwargs["reshape_ensemble_dim"] = False
        out = self.interpolator.predict(torch.cat([initial_condition, x_last], dim=1), condition=static_condition,
                                        time=t, **kwargs)
        return out["preds"]

    # ------------------------------------------------------------------ the sampling loop
    def _refine_times(self) -> List[float]:
        if not self.hparams.refine_intermediate_predictions:
            return []
        steps = self.hparams.prediction_timesteps or list(self.dynamical_steps.values())
        return [i for i in steps if i < self.num_timesteps]

    # ------------------------------------------------------------------ experiment-level inference dropout
    def enable_inference_dropout(self):
        """reference: `enable_inference_dropout(model)` switches EVERY nn.Dropout under the diffusion wrapper to train mode,
        the forecaster's included (src/utilities/utils.py:560-567 via _base_model.py:163-169)."""
        super().enable_inference_dropout()
        for m in (self.model, getattr(self.interpolator, "model", None)):
            if isinstance(m, BaseModel):
                m.enable_inference_dropout()

    def disable_inference_dropout(self):
        super().disable_inference_dropout()
        for m in (self.model, getattr(self.interpolator, "model", None)):
            if isinstance(m, BaseModel):
                m.disable_inference_dropout()

    def _native_ok(self, log_every_t) -> bool:
        ipol = getattr(self.interpolator, "model", None)
        sched = self.sampling_schedule
        return (log_every_t is None and isinstance(self.model, EngineBackbone) and isinstance(ipol, EngineBackbone)
                and sched[-1] == self.num_timesteps - 1 and not self.training)

    def _native_sampler(self, static_channels: int, window_channels: int) -> E.SamplerHandle:
        hp = self.hparams
        sched = list(self.sampling_schedule)
        refine = self._refine_times()
        key = (tuple(sched), tuple(refine), hp.forward_conditioning, hp.sampling_type, hp.time_encoding,
               bool(hp.use_cold_sampling_for_last_step), bool(self.enable_interpolator_dropout), static_channels,
               window_channels, bool(self._inference_dropout))
        F, I = self.model.sync_engine(), self.interpolator.model.sync_engine()
        h = self._native_cache.get(key)
        if h is None:
            tau = [self._sched.interpolation_time(s) if s <= self.num_timesteps - 1 else 0.0 for s in sched]
            h = E.SamplerHandle(
                F, I, num_timesteps=self.num_timesteps, schedule=sched, tau=tau,
                time_forecaster=[self._forecaster_time(s) for s in sched],
                forward_conditioning=hp.forward_conditioning, sampling_type=hp.sampling_type,
                use_cold_sampling_for_last_step=hp.use_cold_sampling_for_last_step, refine_times=refine,
                enable_interpolator_dropout=self.enable_interpolator_dropout, channels=self.num_input_channels,
                window_channels=window_channels, static_channels=static_channels,
                interpolator_horizon=self.interpolator_horizon, max_rows_per_call=hp.get("max_rows_per_call", 0) or 0,
                forecaster_dropout=self._inference_dropout, cuda_graph=hp.get("cuda_graph"))
            self._native_cache[key] = h
        return h

    def sample_loop(self, initial_condition, static_condition: Optional[Tensor] = None,
                    log_every_t: Optional[Union[str, int]] = None, num_predictions: int = None, row_offset: int = 0):
        """`row_offset` (not in the reference): index of row 0 of `initial_condition` in the un-sharded job.  Dropout masks
        and the "data+noise" noise are functions of (seed, call, GLOBAL row), so `distributed.sample_sharded` -- every rank
        holding the same torch seed, as under Lightning's seed_everything -- draws exactly what one rank would draw for the
        whole job, and no two ranks repeat each other's ensemble members."""
        assert len(initial_condition.shape) == 4, f"condition.shape: {initial_condition.shape} (should be 4D)"
        log_every_t = log_every_t or self.hparams.log_every_t
        if not self._native_ok(log_every_t):
            return self._sample_loop_python(initial_condition, static_condition, log_every_t, num_predictions)
        with torch.cuda.device(initial_condition.device):
            sampler = self._native_sampler(0 if static_condition is None else static_condition.shape[1],
                                           initial_condition.shape[1])
        self._calls += 1
        seed = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * self._calls) & 0xFFFFFFFFFFFFFFFF
        preds, x0_hat = sampler.run(initial_condition, static_condition, seed, want_x0=True, row_offset=row_offset)
        out = {}
        for j, k in enumerate(sampler.keys):
            out[f"t{int(k) if float(k).is_integer() else k}_preds"] = preds[j]
        x_s = preds[len(sampler.keys) - 1]
        return x0_hat, out, x_s

    def _sample_loop_python(self, initial_condition, static_condition, log_every_t, num_predictions):
        """Python-driven variant with the reference's exact bookkeeping (dyffusion.py:342-426)."""
        rows, dev = initial_condition.shape[0], initial_condition.device
        log_every_t = 1 if log_every_t == "auto" else log_every_t
        N, sched = self.num_timesteps, list(self.sampling_schedule)
        sc = dict(static_condition=static_condition)
        full = lambda v: torch.full((rows,), v, dtype=torch.float32, device=dev)
        x_s = initial_condition[:, -self.num_input_channels:]
        out, x0_hat, key = {}, None, 0
        after_last = sched[-1] + 1
        x_next = x_cur = None  # x_cur deliberately survives the iteration: on a last step without cold sampling the
        for s, s_next in zip(sched, sched[1:] + [after_last]):  # reference logs the PREVIOUS step's D(x_s, s) (:401-406)
            last = s == N - 1
            x0_hat = self.predict_x_last(condition=initial_condition, x_t=x_s, t=full(s), is_sampling=True, **sc)
            t_next = self.diffusion_step_to_interpolation_step(s_next) if not last else np.inf
            dyn = float(t_next).is_integer() or last
            qkw = dict(x0=x0_hat, x_end=initial_condition, is_artificial_step=not dyn)
            x_next = self.q_sample(**qkw, t=full(s_next), **sc) if s_next <= N - 1 else x0_hat
            if self.hparams.sampling_type == "cold":
                if last and not self.hparams.use_cold_sampling_for_last_step:
                    x_s = x0_hat
                else:
                    x_cur = self.q_sample(**qkw, t=full(s), **sc) if s > 0 else x_s
                    x_s = x_s - x_cur + x_next
            elif self.hparams.sampling_type == "naive":
                x_s = x_next
            else:
                raise ValueError(f"unknown sampling type {self.hparams.sampling_type}")
            key = int(t_next) if s < N - 1 else key + 1
            if dyn:
                out[f"t{key}_preds"] = x_s
                if log_every_t is not None:
                    out[f"t{key}_preds2"] = x_next
            if log_every_t is not None:
                out[f"intermediate_{s}_x0hat"] = x0_hat
                out[f"xipol_{s}_dmodel"] = x_next
                if self.hparams.sampling_type == "cold":
                    out[f"xipol_{s}_dmodel2"] = x_cur
        for i_n in self._refine_times():
            name = int(i_n) if float(i_n).is_integer() else i_n
            assert not float(i_n).is_integer() or f"t{name}_preds" in out, f"t{name}_preds not in intermediates"
            out[f"t{name}_preds"] = self.q_sample(x0=x0_hat, x_end=initial_condition, is_artificial_step=False, t=None,
                                                  interpolation_time=full(i_n), **sc)
        if after_last < N:
            return x_s, out, x_next
        return x0_hat, out, x_s
