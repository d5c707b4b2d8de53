"""Offline checkpoint ingest (first half): a reference Lightning `.ckpt` -> the state dicts of the engine's
drop-in backbones, without wandb or network access.

The reference can only reload checkpoints through its wandb helpers (`src/interface.py:115-172,175-203`); what those do to
the weights is small and is restated here:
  * the Lightning module (`BaseExperiment`) holds the network under `model.`; for a DYffusion run that attribute is the
    diffusion wrapper, so the forecaster backbone sits under `model.model.` and the (frozen) interpolator experiment under
    `model.interpolator.` with its backbone at `model.interpolator.model.` (`src/diffusion/dyffusion.py:461-478`;
    `reload_model_from_config_and_ckpt` drops the `model.interpolator` keys, `interface.py:155-157`);
  * old checkpoints name the linear-attention qkv projection `...fn.to_qkv.weight`; the current module is
    `Sequential(Dropout, Conv2d)`, i.e. `...fn.to_qkv.1.weight`, everywhere except `mid_attn`
    (`rename_state_dict_keys`, `src/utilities/utils.py:530-540`);
  * runs with `use_ema=True` keep an exponential moving average of every trainable parameter of `self.model` in
    `model_ema.<parameter name with the dots removed>` buffers (`LitEma`, `src/models/modules/ema.py:6-27`) and evaluate
    with those weights (`ema_scope`, `_base_experiment.py:263-278`); `use_ema=True` here swaps them in (BatchNorm running
    statistics are buffers, not parameters: they have no shadow and keep their checkpoint values);
  * optimizer state is not a network parameter and is ignored.
The returned dictionaries load with `strict=True` into `dyffusion_b200.backbones.*` (identical state-dict keys, SURVEY.md
A.4).  Pure host bookkeeping: no arithmetic, nothing to run on the GPU.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Union

import torch

Tensors = Dict[str, torch.Tensor]


def rename_state_dict_keys(state_dict: Tensors) -> (Tensors, bool):
    """`src/utilities/utils.py:530-540`: `fn.to_qkv.weight` -> `fn.to_qkv.1.weight` outside `mid_attn`."""
    renamed = False
    for k in list(state_dict.keys()):
        if "fn.to_qkv.weight" in k and "mid_attn" not in k:
            state_dict[k.replace("fn.to_qkv.weight", "fn.to_qkv.1.weight")] = state_dict.pop(k)
            renamed = True
    return state_dict, renamed


def split_state_dict(state_dict: Mapping[str, torch.Tensor], use_ema: bool = False) -> Dict[str, Tensors]:
    """Lightning-module state dict -> {"model": backbone weights[, "interpolator": interpolator-backbone weights]}.

    DYffusion run: forecaster under `model.model.`, interpolator under `model.interpolator.model.`.
    Plain backbone run (e.g. the interpolator's own training run): backbone under `model.`.
    `use_ema`: take the trainable parameters from the `model_ema.*` shadow buffers (what the reference evaluates with when
    the run had `use_ema=True`); ValueError if the checkpoint carries none."""
    sd, _ = rename_state_dict_keys(dict(state_dict))
    out: Dict[str, Tensors] = {}
    interp = {k[len("model.interpolator.model."):]: v for k, v in sd.items() if k.startswith("model.interpolator.model.")}
    if interp:
        out["interpolator"] = interp
    rest = {k: v for k, v in sd.items() if k.startswith("model.") and not k.startswith("model.interpolator.")}
    if any(k.startswith("model.model.") for k in rest):
        inner = "model."  # LitEma walks `self.model` = the diffusion wrapper: parameter names start with its `model.`
        out["model"] = {k[len("model.model."):]: v for k, v in rest.items() if k.startswith("model.model.")}
    else:
        inner = ""
        out["model"] = {k[len("model."):]: v for k, v in rest.items()}
    if not out["model"]:
        raise ValueError("no `model.*` keys: not a state dict of the reference's Lightning modules")
    if use_ema:
        shadows = {k[len("model_ema."):]: v for k, v in sd.items() if k.startswith("model_ema.")}
        shadows.pop("decay", None), shadows.pop("num_updates", None)
        if not shadows:
            raise ValueError("use_ema=True but the checkpoint holds no `model_ema.*` shadow weights (run without use_ema)")
        used = 0
        for k in out["model"]:
            s_name = (inner + k).replace(".", "")  # ema.py:21-24
            if s_name in shadows:
                out["model"][k] = shadows[s_name]
                used += 1
        if used != len(shadows):
            raise ValueError(f"{len(shadows) - used} `model_ema.*` buffers match no backbone parameter")
    return out


def load_reference_checkpoint(path_or_ckpt: Union[str, Mapping], model: Optional[torch.nn.Module] = None,
                              interpolator: Optional[torch.nn.Module] = None, map_location="cpu",
                              use_ema: bool = False) -> Dict[str, object]:
    """Reads a Lightning `.ckpt` (or an already loaded checkpoint / state dict), splits it, and -- if the drop-in backbones
    are given -- loads them strictly.  Returns the split state dicts plus `epoch` / `global_step` when present."""
    ckpt = torch.load(path_or_ckpt, map_location=map_location, weights_only=False) if isinstance(path_or_ckpt, str) else path_or_ckpt
    parts = split_state_dict(ckpt.get("state_dict", ckpt), use_ema=use_ema)
    if model is not None:
        model.load_state_dict(parts["model"], strict=True)
    if interpolator is not None:
        if "interpolator" not in parts:
            raise ValueError("the checkpoint holds no `model.interpolator.model.*` weights")
        target = getattr(interpolator, "model", interpolator)  # an InterpolatorHandle / experiment, or the backbone itself
        target.load_state_dict(parts["interpolator"], strict=True)
    out: Dict[str, object] = dict(parts)
    for k in ("epoch", "global_step"):
        if isinstance(ckpt, Mapping) and k in ckpt:
            out[k] = ckpt[k]
    return out
