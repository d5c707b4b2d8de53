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
  * EMA shadow weights and optimizer state are not network parameters and are ignored.
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


def split_state_dict(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, Tensors]:
    """Lightning-module state dict -> {"model": backbone weights[, "interpolator": interpolator-backbone weights]}.

    DYffusion run: forecaster under `model.model.`, interpolator under `model.interpolator.model.`.
    Plain backbone run (e.g. the interpolator's own training run): backbone under `model.`."""
    sd, _ = rename_state_dict_keys(dict(state_dict))
    out: Dict[str, Tensors] = {}
    interp = {k[len("model.interpolator.model."):]: v for k, v in sd.items() if k.startswith("model.interpolator.model.")}
    if interp:
        out["interpolator"] = interp
    rest = {k: v for k, v in sd.items() if k.startswith("model.") and not k.startswith("model.interpolator.")}
    if any(k.startswith("model.model.") for k in rest):
        out["model"] = {k[len("model.model."):]: v for k, v in rest.items() if k.startswith("model.model.")}
    else:
        out["model"] = {k[len("model."):]: v for k, v in rest.items()}
    if not out["model"]:
        raise ValueError("no `model.*` keys: not a state dict of the reference's Lightning modules")
    return out


def load_reference_checkpoint(path_or_ckpt: Union[str, Mapping], model: Optional[torch.nn.Module] = None,
                              interpolator: Optional[torch.nn.Module] = None, map_location="cpu") -> Dict[str, object]:
    """Reads a Lightning `.ckpt` (or an already loaded checkpoint / state dict), splits it, and -- if the drop-in backbones
    are given -- loads them strictly.  Returns the split state dicts plus `epoch` / `global_step` when present."""
    ckpt = torch.load(path_or_ckpt, map_location=map_location, weights_only=False) if isinstance(path_or_ckpt, str) else path_or_ckpt
    parts = split_state_dict(ckpt.get("state_dict", ckpt))
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
