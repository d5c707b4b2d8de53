"""TEST INFRASTRUCTURE ONLY -- stub packages that let the *unmodified* reference import in this container.

The reference (Rose-STL-Lab/dyffusion, mounted read-only at /root/reference) depends on pytorch_lightning,
omegaconf, hydra, tensordict, torchmetrics, xarray, xskillscore and dask, none of which are installed here.
`install()` registers minimal stand-ins for exactly the members the hot-path import chain touches
 and puts /root/reference on sys.path so `import src...` resolves to the reference.

Used only by `tests/golden/make_golden.py` (fixture generation) and by the `needs_reference` tests that pin
`oracle/` against the real reference when /root/reference is present.  Nothing in the product package
(`dyffusion_b200/`) imports this module, and nothing here runs on the GPU box (no /root/reference there).
"""
from __future__ import annotations

import importlib
import importlib.machinery
import inspect
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DYFFUSION_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "diffusion"))


class AttrDict(dict):
    """dict with attribute access; stands in for omegaconf.DictConfig and Lightning's hparams container."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def get(self, key, default=None, default_value=None):  # omegaconf allows get(key, default_value=...)
        if key in self:
            return super().get(key)
        return default if default is not None else default_value


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)  # torch._dynamo's find_spec scan needs a spec
    m.__path__ = []  # behave as a package so that submodule imports resolve through sys.modules
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _make_lightning():
    import torch
    from torch import nn

    class LightningModule(nn.Module):
        """nn.Module + the handful of Lightning members the reference's hot path uses."""

        def __init__(self, *a, **kw):
            super().__init__()
            self._trainer = None

        # Lightning gathers the ctor args of *every* __init__ frame of the same `self` up the stack.
        def save_hyperparameters(self, *args, ignore=None, **kwargs):
            ignore = set([ignore] if isinstance(ignore, str) else (ignore or []))
            if "_hparams" not in self.__dict__:
                object.__setattr__(self, "_hparams", AttrDict())
            collected = []
            frame = inspect.currentframe().f_back
            while frame is not None:
                loc = frame.f_locals
                if frame.f_code.co_name == "__init__" and loc.get("self") is self:
                    collected.append((frame.f_code, dict(loc)))
                frame = frame.f_back
            for code, loc in collected:
                names = code.co_varnames[: code.co_argcount + code.co_kwonlyargcount]
                for n in names:
                    if n == "self" or n in ignore or n not in loc:
                        continue
                    self._hparams.setdefault(n, loc[n])
                if code.co_flags & inspect.CO_VARKEYWORDS:
                    kwname = code.co_varnames[
                        code.co_argcount + code.co_kwonlyargcount + (1 if code.co_flags & inspect.CO_VARARGS else 0)
                    ]
                    for k, v in (loc.get(kwname) or {}).items():
                        if k not in ignore:
                            self._hparams.setdefault(k, v)

        @property
        def hparams(self):
            if "_hparams" not in self.__dict__:
                object.__setattr__(self, "_hparams", AttrDict())
            return self._hparams

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        @property
        def trainer(self):
            return self._trainer

        @trainer.setter
        def trainer(self, t):
            object.__setattr__(self, "_trainer", t)

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    def rank_zero_only(fn):
        return fn

    pl = _mod(
        "pytorch_lightning",
        LightningModule=LightningModule,
        LightningDataModule=_Dummy,
        Callback=_Dummy,
        Trainer=_Dummy,
        seed_everything=lambda seed=0, **k: torch.manual_seed(seed),
    )
    util = _mod("pytorch_lightning.utilities", rank_zero_only=rank_zero_only)
    types_ = _mod("pytorch_lightning.utilities.types", EVAL_DATALOADERS=object, TRAIN_DATALOADERS=object)
    cbs = _mod("pytorch_lightning.callbacks", ModelCheckpoint=_Dummy, Callback=_Dummy)
    loggers = _mod("pytorch_lightning.loggers", WandbLogger=_Dummy)
    _mod("pytorch_lightning.loggers.wandb", WandbLogger=_Dummy)
    pl.utilities, util.types, pl.callbacks, pl.loggers = util, types_, cbs, loggers


def _instantiate(cfg, *args, _recursive_=False, **kwargs):
    """hydra.utils.instantiate stand-in: import `_target_`, call with {**cfg, **kwargs}; objects pass through."""
    merged = {k: v for k, v in dict(cfg).items() if k not in ("_target_", "_recursive_")}
    merged.update(kwargs)
    module_name, _, cls_name = cfg["_target_"].rpartition(".")
    cls = getattr(importlib.import_module(module_name), cls_name)
    return cls(*args, **merged)


def install() -> None:
    """Idempotently register the stubs and expose the reference as the top-level package `src`."""
    if getattr(install, "_done", False):
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _make_lightning()

    class _OmegaConf:
        @staticmethod
        def create(x=None):
            return AttrDict(x or {})

        @staticmethod
        def to_container(x, **k):
            return dict(x)

        @staticmethod
        def set_struct(*a, **k):
            pass

    class _open_dict:
        def __init__(self, cfg):
            self.cfg = cfg

        def __enter__(self):
            return self.cfg

        def __exit__(self, *a):
            return False

    _mod("omegaconf", DictConfig=AttrDict, OmegaConf=_OmegaConf, open_dict=_open_dict, ListConfig=list)
    hy = _mod("hydra")
    hy.utils = _mod("hydra.utils", instantiate=_instantiate)
    _mod("tensordict", TensorDict=dict)

    class _Metric:
        def __init__(self, *a, **k):
            pass

    _mod("torchmetrics", MeanSquaredError=_Metric, Metric=_Metric)
    _mod("xarray", DataArray=object, Dataset=object)
    _mod("xskillscore")
    _mod("dask")
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)  # at the END: the reference has its own (empty) `tests` package
    install._done = True
