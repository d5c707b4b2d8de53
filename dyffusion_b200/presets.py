"""The reference's three shipped DYffusion configurations as ready-made engine objects.

The hyper-parameters live in `dyffusion_b200/configs/` as files: `model/*_b200.yaml` and `diffusion/dyffusion_b200.yaml` are
Hydra configs in the reference's own layout (copy them into `src/configs/` and select them with `model=... diffusion=...`,
INTEGRATION.md); `experiment/*_b200.yaml` flatten the per-dataset overrides of the reference's experiment files.  This module
reads them without Hydra and instantiates the drop-in classes the way `BaseExperiment.instantiate_model`
(src/experiment_types/_base_experiment.py:137-196) and `InterpolationExperiment` / `MultiHorizonForecastingDYffusion`
would: channel bookkeeping of SURVEY.md A.5 (forecaster input = C, condition = static [+ window*C unless
forward_conditioning == "none"]; interpolator input = window*C + C, condition = static)."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional, Tuple

import torch
import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")
PRESETS = {"ns": "navier_stokes_dyffusion_b200.yaml", "sst": "oisst_pacific_dyffusion_b200.yaml",
           "spring": "spring_mesh_dyffusion_b200.yaml"}
_NOT_CTOR_KEYS = ("_target_",)


def _load(rel: str) -> Dict[str, Any]:
    with open(os.path.join(CONFIG_DIR, rel)) as f:
        return yaml.safe_load(f)


def load_preset(name: str) -> Dict[str, Any]:
    """{'dataset', 'model', 'interpolator_model', 'diffusion', 'evaluation', 'model_target', 'diffusion_target'} with the
    Hydra model / diffusion defaults merged under the experiment's overrides."""
    if name not in PRESETS:
        raise ValueError(f"unknown preset {name!r}; choose from {sorted(PRESETS)}")
    exp = _load(os.path.join("experiment", PRESETS[name]))
    mc = _load(exp["model_config"])
    mc = mc.get("model", mc)  # cnn_simple.yaml is not `@package _global_`
    model = {k: v for k, v in mc.items() if k not in ("defaults", "trainer") + _NOT_CTOR_KEYS}
    model.update(exp.get("model") or {})
    diff = dict(_load(os.path.join("diffusion", "dyffusion_b200.yaml"))["diffusion"])
    diff_target = diff.pop("_target_")
    for k in ("interpolator", "interpolator_run_id", "interpolator_wandb_ckpt_filename",
              "interpolator_local_checkpoint_path"):
        diff.pop(k, None)
    diff.update(exp.get("diffusion") or {})
    diff["timesteps"] = exp["dataset"]["horizon"]  # `${datamodule.horizon}`
    return dict(dataset=exp["dataset"], model=model, interpolator_model=dict(model, **(exp.get("interpolator_model") or {})),
                diffusion=diff, evaluation=exp.get("evaluation") or {}, model_target=mc["_target_"],
                diffusion_target=diff_target)


def _resolve(target: str):
    import importlib
    mod, cls = target.rsplit(".", 1)
    return getattr(importlib.import_module(mod), cls)


def channels(preset: Dict[str, Any], role: str) -> Tuple[int, int, int]:
    """(num_input_channels, num_conditional_channels, num_output_channels) of the forecaster ("F") / interpolator ("I")."""
    d = preset["dataset"]
    c, st, w = d["channels"], d["static_channels"], d["window"]
    if role == "I":
        return c * w + c, st, c
    extra = 0 if preset["diffusion"]["forward_conditioning"] in ("none", None, "") else w * c
    return c, st + extra, c


def build_backbone(preset: Dict[str, Any], role: str, device="cuda"):
    cin, ccond, cout = channels(preset, role)
    kw = dict(preset["interpolator_model" if role == "I" else "model"])
    cls = _resolve(preset["model_target"])
    net = cls(**kw, num_input_channels=cin, num_output_channels=cout, num_conditional_channels=ccond,
              spatial_shape=tuple(preset["dataset"]["spatial_shape"]), verbose=False)
    return net.to(device).eval()


def randomize_norm_statistics(module: torch.nn.Module, seed: int = 0) -> None:
    """Non-trivial BatchNorm running statistics for synthetic-weight runs (defaults of mean 0 / var 1 would make the folded
    normalisation an identity): running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, buf in module.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))


def build_dyffusion(name: str, device="cuda", seed: Optional[int] = 0, **diffusion_overrides):
    """The DYffusion drop-in of a shipped configuration with freshly initialised (reference default init, `seed`) weights:
    load trained ones with `load_state_dict` / `dyffusion_b200.checkpoint`."""
    from .diffusion import InterpolatorHandle

    preset = load_preset(name)
    preset["diffusion"].update(diffusion_overrides)
    if seed is not None:
        torch.manual_seed(seed)
    F = build_backbone(preset, "F", device)
    I = build_backbone(preset, "I", device)
    if seed is not None:
        randomize_norm_statistics(F, seed + 1)
        randomize_norm_statistics(I, seed + 2)
    ipol = InterpolatorHandle(I, horizon=preset["dataset"]["horizon"], window=preset["dataset"]["window"])
    cls = _resolve(preset["diffusion_target"])
    return cls(model=F, interpolator=ipol, verbose=False, **preset["diffusion"]).to(device).eval()
