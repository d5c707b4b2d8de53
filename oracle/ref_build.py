"""TEST INFRASTRUCTURE -- builds the *real* reference modules (imported unmodified from /root/reference through
`oracle/ref_shims.py`) at the configs in `oracle/configs.py`.  Only usable where /root/reference exists (the build
container); used by `tests/golden/make_golden.py` and the live oracle-vs-reference tests."""
from __future__ import annotations

import zlib
from copy import deepcopy

import torch
from torch import nn

from . import configs as C
from . import ref_shims
from .synth import _gen


def _cfg(d):
    return ref_shims.AttrDict(deepcopy(d))


def build_interpolator(dataset: str, horizon: int, model_overrides=None):
    """reference InterpolationExperiment (src/experiment_types/interpolation.py:12-167) + its backbone."""
    ref_shims.install()
    from src.experiment_types.interpolation import InterpolationExperiment

    m = C.MODELS[dataset]
    mk = dict(m["kwargs"], **C.INTERPOLATOR_OVERRIDES[dataset], **(model_overrides or {}))
    mc = _cfg(dict(_target_=m["ref_target"], name="", verbose=False, loss_function="mse", **mk))
    dm = _cfg(dict(C.DATASETS[dataset]["datamodule"], horizon=horizon))
    exp = InterpolationExperiment(model_config=mc, datamodule_config=dm, verbose=False, num_predictions=1)
    return exp.eval()


def build_dyffusion(dataset: str, interpolator, model_overrides=None, **diffusion_overrides):
    """reference MultiHorizonForecastingDYffusion (src/experiment_types/forecasting_multi_horizon.py:398-424)
    wrapping DYffusion (src/diffusion/dyffusion.py:439-567) around the forecaster backbone."""
    ref_shims.install()
    from src.experiment_types.forecasting_multi_horizon import MultiHorizonForecastingDYffusion

    m = C.MODELS[dataset]
    mk = dict(m["kwargs"], **(model_overrides or {}))
    mc = _cfg(dict(_target_=m["ref_target"], name="", verbose=False, loss_function="mse", **mk))
    dk = C.diffusion_kwargs(dataset, **diffusion_overrides)
    dm = _cfg(dict(C.DATASETS[dataset]["datamodule"], horizon=dk["timesteps"]))
    dc = _cfg(dict(_target_="src.diffusion.dyffusion.DYffusion", **dk))
    dc["interpolator"] = interpolator  # the live module must pass through un-copied (interface.py:182-186)
    exp = MultiHorizonForecastingDYffusion(model_config=mc, datamodule_config=dm, diffusion_config=dc, verbose=False,
                                           num_predictions=1)
    return exp.eval()


def dropout_site(module_name: str) -> str:
    """reference nn.Dropout module path -> oracle dropout-site name."""
    n = module_name
    if n.endswith(".fn.fn.to_qkv.0"):
        return n[: -len(".fn.fn.to_qkv.0")] + ".to_qkv"
    if n.endswith(".fn.fn.dropout"):
        return n[: -len(".fn.fn.dropout")] + ".attn"
    if n.endswith(".dropout"):
        return n[: -len(".dropout")]
    return n


class HookedDropout:
    """Replaces every nn.Dropout of a reference backbone by the deterministic site masks of synth.SiteDropout."""

    def __init__(self, backbone: nn.Module, seed: int = 0):
        self.seed, self.calls, self.handles = seed, {}, []
        for name, mod in backbone.named_modules():
            if isinstance(mod, nn.Dropout):
                self.handles.append(mod.register_forward_hook(self._make(dropout_site(name), mod)))

    def _make(self, site, mod):
        def hook(_m, inputs, _out):
            x, p = inputs[0], mod.p
            if p <= 0:
                return x
            n = self.calls.get(site, 0)
            self.calls[site] = n + 1
            keep = torch.rand(x.shape, generator=_gen(f"{site}#{n}", self.seed)) >= p
            return x * keep.to(x.dtype) / (1.0 - p)

        return hook

    def remove(self):
        for h in self.handles:
            h.remove()
