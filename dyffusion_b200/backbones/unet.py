"""Drop-in for `src.models.unet.Unet` (reference src/models/unet.py:114-315), the SST backbone: weight-standardised
3x3 convs + GroupNorm + SiLU ResNet blocks, linear attention per level, full attention in the bottleneck."""
from __future__ import annotations

from .. import engine as E
from .._base import EngineBackbone


class Unet(EngineBackbone):
    def __init__(self, dim, init_dim=None, dim_mults=(1, 2, 4, 8), num_conditions: int = 0, resnet_block_groups=8,
                 with_time_emb: bool = False, block_dropout: float = 0.0, block_dropout1: float = 0.0,
                 attn_dropout: float = 0.0, input_dropout: float = 0.0, double_conv_layer: bool = True,
                 learned_variance=False, learned_sinusoidal_cond=False, learned_sinusoidal_dim=16,
                 outer_sample_mode: str = None, upsample_dims: tuple = None, keep_spatial_dims: bool = False,
                 init_kernel_size: int = 7, init_padding: int = 3, init_stride: int = 1, **kwargs):
        for flag, name in ((learned_variance, "learned_variance"), (learned_sinusoidal_cond, "learned_sinusoidal_cond"),
                           (not double_conv_layer, "double_conv_layer=False"), (outer_sample_mode, "outer_sample_mode"),
                           (upsample_dims, "upsample_dims"), (init_dim not in (None, dim), "init_dim != dim")):
            if flag:
                raise NotImplementedError(f"Unet option {name} is not built")
        if len(dim_mults) > 8:
            raise NotImplementedError("at most 8 resolutions")
        h, w = kwargs["spatial_shape"]
        d = E.NetDesc(arch=E.ARCH_UNET_RESNET, dim=dim, in_channels=kwargs["num_input_channels"],
                      cond_channels=kwargs.get("num_conditional_channels", 0) or 0,
                      out_channels=kwargs["num_output_channels"], height=h, width=w, with_time_emb=int(with_time_emb),
                      input_dropout=float(input_dropout), n_mults=len(dim_mults), groups=int(resnet_block_groups),
                      block_dropout=float(block_dropout), block_dropout1=float(block_dropout1),
                      attn_dropout=float(attn_dropout), keep_spatial_dims=int(bool(keep_spatial_dims)),
                      init_kernel=int(init_kernel_size), init_padding=int(init_padding), init_stride=int(init_stride))
        for i, m in enumerate(dim_mults):
            d.dim_mults[i] = int(m)
        super().__init__(d, **kwargs)
        loc = dict(locals())
        for k in ("self", "kwargs", "d", "h", "w", "i", "m", "flag", "name", "__class__"):
            loc.pop(k, None)
        self._record_hparams(loc)
        self.time_dim = dim * 2 if with_time_emb else None
