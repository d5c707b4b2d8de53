"""Host-side mirror of the reference's `BaseModel` surface (src/models/_base_model.py:19-175) for engine-backed
backbones: same attributes and methods the experiment/diffusion classes touch, with `forward` routed to the
C-ABI engine instead of torch ops."""
from __future__ import annotations

import logging
from contextlib import contextmanager
from typing import Any, Dict, Optional, Sequence, Union

import torch
from torch import Tensor, nn

from . import engine as E

try:  # the reference builds on Lightning; when it is installed the drop-ins are LightningModules too
    from pytorch_lightning import LightningModule as _ModuleBase  # type: ignore
    _HAVE_LIGHTNING = hasattr(_ModuleBase, "save_hyperparameters") and hasattr(_ModuleBase, "log_dict")
except Exception:  # pragma: no cover - lightning is absent in the build image
    _ModuleBase, _HAVE_LIGHTNING = nn.Module, False


class HParams(dict):
    """attribute-dict standing in for Lightning's `hparams` (supports .get/in/pop and item/attr assignment)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def get_loss(name: str):
    """src/utilities/utils.py:201-212."""
    n = name.lower().strip().replace("-", "_")
    if n in ("l1", "mae", "mean_absolute_error"):
        return nn.L1Loss()
    if n in ("l2", "mse", "mean_squared_error"):
        return nn.MSELoss()
    if n in ("smoothl1", "smooth"):
        return nn.SmoothL1Loss()
    raise ValueError(f"Unknown loss function {name}")


class _Holder(nn.Module):
    """Anonymous container used to reproduce the reference's dotted state-dict keys."""


class EngineModule(_ModuleBase):
    """nn.Module with the hparams / dropout-scope plumbing shared by backbones and the diffusion wrapper."""

    def __init__(self):
        super().__init__()
        if not _HAVE_LIGHTNING and "_hp" not in self.__dict__:
            object.__setattr__(self, "_hp", HParams())

    if not _HAVE_LIGHTNING:
        @property
        def hparams(self) -> HParams:
            return self._hp

        @property
        def device(self) -> torch.device:
            for t in list(self.parameters(recurse=True))[:1] + list(self.buffers(recurse=True))[:1]:
                return t.device
            return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    def _record_hparams(self, values: Dict[str, Any], ignore: Sequence[str] = ()) -> None:
        for k, v in values.items():
            if k in ("self", "__class__", "kwargs", "args") or k in ignore or k.startswith("_"):
                continue
            if k not in self.hparams:
                self.hparams[k] = v


class BaseModel(EngineModule):
    """Mirror of src/models/_base_model.py:BaseModel (constructor :42-75)."""

    def __init__(self, num_input_channels: int = None, num_output_channels: int = None,
                 num_conditional_channels: int = 0, spatial_shape: Union[Sequence[int], int] = None,
                 loss_function: str = "mean_squared_error", datamodule_config: Optional[Any] = None, name: str = "",
                 verbose: bool = True):
        super().__init__()
        self._record_hparams(dict(num_input_channels=num_input_channels, num_output_channels=num_output_channels,
                                  num_conditional_channels=num_conditional_channels, spatial_shape=spatial_shape,
                                  loss_function=loss_function, datamodule_config=datamodule_config, name=name))
        self.log_text = logging.getLogger(self.__class__.__name__ if name == "" else name)
        self.name, self.verbose = name, verbose
        if not verbose:
            self.log_text.setLevel(logging.WARN)
        self.num_input_channels = num_input_channels
        self.num_output_channels = num_output_channels
        self.num_conditional_channels = num_conditional_channels
        self.spatial_shape = spatial_shape
        self.datamodule_config = datamodule_config
        self.criterion = get_loss(loss_function)
        self._channel_dim = None
        self.ema_scope = None
        self._inference_dropout = False

    # ---- `ema_scope` is assigned by the reference's BaseExperiment (src/experiment_types/_base_experiment.py:107).  Its
    # LitEma swaps weights with `param.data.copy_()` (src/models/modules/ema.py:48-78), which no version counter sees, so the
    # EMA object's copy_to / restore are wrapped to flag the engine's packed weights as stale.
    @property
    def ema_scope(self):
        return self.__dict__.get("_ema_scope")

    @ema_scope.setter
    def ema_scope(self, fn):
        self.__dict__["_ema_scope"] = fn
        ema = getattr(getattr(fn, "__self__", None), "model_ema", None)
        if ema is None or getattr(ema, "_dyf_hooked", False):
            return
        for name in ("copy_to", "restore"):
            orig = getattr(ema, name, None)
            if orig is None:
                continue

            def hooked(*a, _orig=orig, **k):
                out = _orig(*a, **k)
                self.mark_engine_dirty()
                return out

            object.__setattr__(ema, name, hooked)
        object.__setattr__(ema, "_dyf_hooked", True)

    def mark_engine_dirty(self) -> None:
        """Force a re-pack of every engine backbone under this module before its next call.  Needed only after writes that
        bypass autograd's version counters (`param.data.copy_()`, `param.data = ...` keeping the storage); optimizer steps,
        nn.init, load_state_dict, .to()/.half() and the reference's EMA scope are detected automatically."""
        for m in self.modules():
            if hasattr(m, "_dirty"):
                m._dirty = True

    # ---- reference surface (:77-175)
    @property
    def short_description(self) -> str:
        return self.name if self.name else self.__class__.__name__

    def get_parameters(self) -> list:
        return list(self.parameters())

    @property
    def num_params(self):
        return sum(p.numel() for p in self.get_parameters() if p.requires_grad)

    @property
    def channel_dim(self):
        return 1

    def get_loss(self, inputs: Tensor, targets: Tensor, condition: Tensor = None, metadata: Any = None,
                 predictions_mask: Optional[Tensor] = None, return_predictions: bool = False, **kwargs):
        predictions = self(inputs, condition=condition, **kwargs)
        loss = self.criterion(predictions[predictions_mask] if predictions_mask is not None else predictions, targets)
        return (loss, predictions) if return_predictions else loss

    def predict_forward(self, inputs: Tensor, metadata: Any = None, **kwargs):
        return self(inputs, **kwargs)

    @contextmanager
    def inference_dropout_scope(self, condition: bool, context=None):
        assert isinstance(condition, bool), f"Condition must be a boolean, got {condition}"
        if condition:
            self.enable_inference_dropout()
        try:
            yield None
        finally:
            if condition:
                self.disable_inference_dropout()

    def enable_inference_dropout(self):
        """reference: set all nn.Dropout layers to train mode (src/utilities/utils.py:560-567)."""
        self._inference_dropout = True

    def disable_inference_dropout(self):
        self._inference_dropout = False


class EngineBackbone(BaseModel):
    """A backbone whose parameters live in an nn.Module tree with the reference's state-dict keys and whose
    forward runs in the CUDA engine."""

    def __init__(self, desc: E.NetDesc, **base_kwargs):
        super().__init__(**base_kwargs)
        self._net = E.NetHandle(desc)
        self._dirty = True
        self._drop_stream = 0
        self._specs = self._net.param_specs()
        self._spec_keys = {s[0] for s in self._specs}
        self._packed_fp = None
        for key, shape, is_buffer in self._specs:
            self._register(key, shape, is_buffer)
        self.reset_parameters()

    # ---- parameter tree with reference keys
    def _register(self, key: str, shape, is_buffer: bool) -> None:
        *path, leaf = key.split(".")
        mod: nn.Module = self
        for part in path:
            if part not in mod._modules:
                mod.add_module(part, _Holder())
            mod = mod._modules[part]
        if leaf == "num_batches_tracked":
            mod.register_buffer(leaf, torch.zeros((), dtype=torch.long))
        elif is_buffer:
            mod.register_buffer(leaf, torch.zeros(shape))
        else:
            mod.register_parameter(leaf, nn.Parameter(torch.zeros(shape)))

    def reset_parameters(self) -> None:
        """Default initialisation: PyTorch's conv/linear defaults, unit norms; subclasses refine it."""
        sd = self.state_dict()
        with torch.no_grad():
            for key, shape, _ in self._specs:
                t, leaf = sd[key], key.rsplit(".", 1)[-1]
                if leaf == "running_var":
                    t.fill_(1.0)
                elif leaf in ("running_mean", "num_batches_tracked"):
                    t.zero_()
                elif leaf == "g":
                    t.fill_(1.0)
                elif t.dim() >= 2:
                    fan_in = t[0].numel()
                    bound = (1.0 / fan_in) ** 0.5
                    t.uniform_(-bound, bound)
                elif leaf == "weight":
                    t.fill_(1.0)
                else:
                    t.zero_()
        self._dirty = True

    # ---- keep the engine's packed weights in sync with the nn.Parameters
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._dirty = True
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._dirty = True
        return out

    def mark_dirty(self) -> None:
        self._dirty = True

    def _fingerprint(self, own) -> tuple:
        """(storage pointer, in-place version counter) of every tensor the engine packed: any in-place write -- an optimizer
        step, `p.data.copy_()`, nn.init, the reference's `ema_scope` (LitEma.copy_to / restore, src/models/modules/ema.py) --
        bumps `_version`; a re-assigned `.data` changes the pointer."""
        return tuple((v.data_ptr(), v._version) for v in own.values())

    def sync_engine(self) -> E.NetHandle:
        """Re-pack the engine's weights whenever the nn.Parameters / buffers changed since the last pack."""
        keys = self._spec_keys
        own = {k: v for k, v in self.state_dict(keep_vars=True).items() if k in keys}
        fp = self._fingerprint(own)
        if self._dirty or not self._net.finalized or fp != self._packed_fp:
            dev = next(iter(own.values())).device
            if dev.type != "cuda":
                raise E.EngineError(f"dyffusion_b200 has no CPU path: move the module to a CUDA device first (parameters on {dev})")
            with torch.cuda.device(dev):  # pack on the GPU that owns the parameters, whatever the current device is
                self._net.load({k: v.detach() for k, v in own.items()})
            self._dirty = False
            self._packed_fp = fp
        return self._net

    def _dropout_arg(self, row_offset: int = 0):
        if not (self._inference_dropout or self.training):
            return None
        self._drop_stream += 1
        return (int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF, self._drop_stream, int(row_offset))

    def forward(self, inputs, time=None, condition=None, return_time_emb: bool = False, row_offset: int = 0, **kwargs):
        """`row_offset`: index of row 0 in the un-sharded job (dropout masks are functions of the global row)."""
        if return_time_emb:
            raise NotImplementedError("return_time_emb=True is not exposed by the engine")
        if self.num_conditional_channels == 0 and condition is not None:
            raise AssertionError("condition is not None but num_conditional_channels is 0")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("training through the CUDA engine is not built yet (SURVEY.md 8f-1); "
                                      "use torch.no_grad()/eval() for sampling")
        net = self.sync_engine()
        return net.forward(inputs, time, condition, dropout=self._dropout_arg(row_offset))
