"""Builders for the engine-backed drop-in classes at the oracle/configs.py settings (test infrastructure)."""
from __future__ import annotations

import torch

from oracle import configs as C
from oracle.synth import synth_state_dict
from tests import helpers as H


def build_backbone(dataset, role, device="cuda", seed=None, fcond=None, **overrides):
    from dyffusion_b200.backbones import SimpleConvNet, UNet, Unet

    kw = H.model_kwargs(dataset, role)
    kw.update(overrides)
    fcond = fcond if fcond is not None else C.DIFFUSION[dataset]["forward_conditioning"]
    cin, ccond, cout = C.channels(dataset, role, fcond)
    base = dict(num_input_channels=cin, num_output_channels=cout, num_conditional_channels=ccond,
                spatial_shape=C.DATASETS[dataset]["spatial"], verbose=False)
    cls = {"unet_simple": UNet, "unet_resnet": Unet, "simple_conv_net": SimpleConvNet}[C.MODELS[dataset]["arch"]]
    m = cls(**kw, **base)
    if seed is not None:
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(synth_state_dict(shapes, seed=seed), strict=True)
    return m.to(device).eval()


def build_dyffusion(dataset, device="cuda", seeds=(3, 2), interpolator_horizon=None, backbone_arch_override=None,
                    **diffusion_overrides):
    """(forecaster seed, interpolator seed) follow tests/golden/make_golden.py."""
    from dyffusion_b200.diffusion import DYffusion, InterpolatorHandle

    dk = C.diffusion_kwargs(dataset, **diffusion_overrides)
    ds_model = backbone_arch_override or dataset
    fcond = dk["forward_conditioning"]
    if backbone_arch_override:  # host-logic tests only: borrow another dataset's backbone for the schedule checks
        F = build_backbone(ds_model, "F", device, seed=None, fcond=fcond)
        I = build_backbone(ds_model, "I", device, seed=None)
    else:
        F = build_backbone(dataset, "F", device, seed=seeds[0] if device != "cpu" else None, fcond=fcond)
        I = build_backbone(dataset, "I", device, seed=seeds[1] if device != "cpu" else None)
    ipol = InterpolatorHandle(I, horizon=interpolator_horizon or dk["timesteps"])
    return DYffusion(model=F, interpolator=ipol, verbose=False, **dk).to(device).eval()
