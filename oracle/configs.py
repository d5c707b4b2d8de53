"""TEST INFRASTRUCTURE -- the reference's shipped experiment settings for the three datasets on the hot path,
as plain dicts (sources: /root/reference/src/configs/experiment/{navier_stokes,oisst_pacific,spring_mesh}*.yaml,
model/{unet_simple_navier_stokes,unet_resnet,cnn_simple}.yaml, diffusion/dyffusion.yaml,
src/datamodules/dataset_dimensions.py:4-22; channel bookkeeping SURVEY.md A.5)."""
from __future__ import annotations

from copy import deepcopy

DATASETS = {
    # name: (C, C_static_cond, (H, W), datamodule `_target_`/fields the reference's get_dims_of_dataset keys on)
    "ns": dict(channels=3, static=2, spatial=(221, 42),
               datamodule=dict(_target_="src.datamodules.physical_systems_benchmark.PhysicalSystemsBenchmarkDataModule",
                               physical_system="navier-stokes", window=1)),
    "sst": dict(channels=1, static=0, spatial=(60, 60),
                datamodule=dict(_target_="src.datamodules.oisstv2.OISSTv2DataModule", box_size=60, window=1)),
    "spring": dict(channels=4, static=1, spatial=(10, 10),
                   datamodule=dict(_target_="src.datamodules.physical_systems_benchmark.PhysicalSystemsBenchmarkDataModule",
                                   physical_system="spring-mesh", window=1)),
}

MODELS = {
    "ns": dict(arch="unet_simple", ref_target="src.models.unet_simple.UNet",
               kwargs=dict(dim=64, with_time_emb=True, outer_sample_mode="bilinear", upsample_dims=[256, 256],
                           dropout=0.15, input_dropout=0.0)),
    "sst": dict(arch="unet_resnet", ref_target="src.models.unet.Unet",
                kwargs=dict(dim=64, dim_mults=[1, 2, 4], resnet_block_groups=8, double_conv_layer=True,
                            learned_variance=False, learned_sinusoidal_cond=False, learned_sinusoidal_dim=16,
                            input_dropout=0.0, block_dropout=0.3, block_dropout1=0.0, attn_dropout=0.1,
                            with_time_emb=True, keep_spatial_dims=False, outer_sample_mode=None, upsample_dims=None,
                            init_kernel_size=7, init_padding=3, init_stride=1)),
    "spring": dict(arch="simple_conv_net", ref_target="src.models.simple_conv_net.SimpleConvNet",
                   kwargs=dict(dim=64, kernel_sizes=[9, 7, 5, 3], residual=True, net_normalization="batch_norm",
                               dropout=0.05, with_time_emb=True)),
}
# interpolator-side dropout overrides (experiment/*_interpolation.yaml)
INTERPOLATOR_OVERRIDES = {"ns": {}, "sst": dict(block_dropout=0.6, block_dropout1=0.2, attn_dropout=0.6), "spring": {}}

DIFFUSION_DEFAULTS = dict(  # diffusion/dyffusion.yaml
    lambda_reconstruction=0.5, lambda_reconstruction2=0.5, forward_conditioning="data", schedule="before_t1_only",
    additional_interpolation_steps=0, additional_interpolation_steps_factor=0, interpolate_before_t1=True,
    time_encoding="dynamics", enable_interpolator_dropout=True, sampling_type="cold", sampling_schedule=None,
    refine_intermediate_predictions=False, use_cold_sampling_for_last_step=False, log_every_t=None)

DIFFUSION = {  # experiment/*_dyffusion.yaml
    "ns": dict(horizon=16, refine_intermediate_predictions=True, forward_conditioning="none"),
    "sst": dict(horizon=7, additional_interpolation_steps=25, refine_intermediate_predictions=False,
                forward_conditioning="data+noise"),
    "spring": dict(horizon=134, refine_intermediate_predictions=True, forward_conditioning="data"),
}


def channels(dataset: str, role: str, forward_conditioning: str, window: int = 1):
    """(C_in, C_cond, C_out) of the forecaster ('F') / interpolator ('I') -- SURVEY.md A.5."""
    d = DATASETS[dataset]
    c, st = d["channels"], d["static"]
    if role == "I":
        return c * window + c, st, c
    extra = 0 if forward_conditioning in ("none", None, "") else window * c
    return c, st + extra, c


def diffusion_kwargs(dataset: str, **overrides) -> dict:
    kw = deepcopy(DIFFUSION_DEFAULTS)
    spec = deepcopy(DIFFUSION[dataset])
    spec.update(overrides)
    horizon = spec.pop("horizon")
    kw.update(spec)
    kw["timesteps"] = horizon
    return kw
