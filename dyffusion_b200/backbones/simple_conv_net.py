"""Drop-in for `src.models.simple_conv_net.SimpleConvNet` (reference src/models/simple_conv_net.py:59-131), the
spring-mesh backbone."""
from __future__ import annotations

from typing import Sequence

from .. import engine as E
from .._base import EngineBackbone


class SimpleConvNet(EngineBackbone):
    def __init__(self, dim: int, with_time_emb: bool = False, net_normalization: str = "batch_norm",
                 kernel_sizes: Sequence[int] = (7, 3, 3), keep_spatial_shape: bool = True, residual=True,
                 dropout: float = 0.0, *args, **kwargs):
        if net_normalization != "batch_norm":
            raise NotImplementedError(f"net_normalization={net_normalization!r}: only 'batch_norm' is built")
        if not keep_spatial_shape:
            raise NotImplementedError("keep_spatial_shape=False is not built")
        if len(kernel_sizes) > 8:
            raise NotImplementedError("at most 8 conv blocks")
        h, w = kwargs["spatial_shape"]
        d = E.NetDesc(arch=E.ARCH_CONVNET, dim=dim, in_channels=kwargs["num_input_channels"],
                      cond_channels=kwargs.get("num_conditional_channels", 0) or 0,
                      out_channels=kwargs["num_output_channels"], height=h, width=w, with_time_emb=int(with_time_emb),
                      dropout=float(dropout), n_kernels=len(kernel_sizes), residual=int(bool(residual)))
        for i, k in enumerate(kernel_sizes):
            d.kernel_sizes[i] = int(k)
        super().__init__(d, *args, **kwargs)
        self._record_hparams(dict(dim=dim, with_time_emb=with_time_emb, net_normalization=net_normalization,
                                  kernel_sizes=kernel_sizes, keep_spatial_shape=keep_spatial_shape, residual=residual,
                                  dropout=dropout))
        self.time_dim = dim * 2 if with_time_emb else None
