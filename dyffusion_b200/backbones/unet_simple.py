"""Drop-in for `src.models.unet_simple.UNet` (reference src/models/unet_simple.py:86-197), the Navier-Stokes
backbone: same constructor arguments and state-dict keys; the forward pass is the engine's kernel sequence
(resize+pack -> implicit-GEMM convs with fused BN/time/activation/dropout epilogues -> fused readout)."""
from __future__ import annotations

import torch

from .. import engine as E
from .._base import EngineBackbone


class UNet(EngineBackbone):
    def __init__(self, dim: int, with_time_emb: bool = False, outer_sample_mode: str = "bilinear",
                 upsample_dims: tuple = (256, 256), dropout: float = 0.0, input_dropout: float = 0.0, **kwargs):
        if upsample_dims is not None and outer_sample_mode != "bilinear":
            raise NotImplementedError(f"outer_sample_mode={outer_sample_mode!r}: only 'bilinear' is built")
        h, w = kwargs["spatial_shape"]
        d = E.NetDesc(arch=E.ARCH_UNET_SIMPLE, dim=dim, in_channels=kwargs["num_input_channels"],
                      cond_channels=kwargs.get("num_conditional_channels", 0) or 0,
                      out_channels=kwargs["num_output_channels"], height=h, width=w, with_time_emb=int(with_time_emb),
                      upsample_h=0 if upsample_dims is None else int(upsample_dims[0]),
                      upsample_w=0 if upsample_dims is None else int(upsample_dims[1]),
                      dropout=float(dropout), input_dropout=float(input_dropout))
        self.outer_sample_mode = outer_sample_mode
        super().__init__(d, **kwargs)
        self._record_hparams(dict(dim=dim, with_time_emb=with_time_emb, outer_sample_mode=outer_sample_mode,
                                  upsample_dims=upsample_dims, dropout=dropout, input_dropout=input_dropout))
        self.time_dim = dim * 2 if with_time_emb else None

    def reset_parameters(self) -> None:
        """reference init (unet_simple.py:156-162): conv weights N(0, 0.02), BatchNorm weight N(1, 0.02), bias 0."""
        super().reset_parameters()
        sd = self.state_dict()
        with torch.no_grad():
            for key, shape, _ in self._specs:
                if len(shape) == 4:
                    sd[key].normal_(0.0, 0.02)
                elif key.endswith(".weight") and len(shape) == 1 and key.replace(".weight", ".running_mean") in sd:
                    sd[key].normal_(1.0, 0.02)
