"""Drop-in check against the REAL reference experiment classes (imported through the shims; build container only):
the reference's `InterpolationExperiment` / `MultiHorizonForecastingDYffusion` instantiate our `_target_`s through
`hydra.utils.instantiate` exactly as `run.py` would (src/experiment_types/_base_experiment.py:173-199), and the
resulting module tree exposes what the reference's evaluation code reads.  No arithmetic runs here (no GPU)."""
import pytest
import torch

from oracle import configs as C
from oracle import ref_shims

pytestmark = pytest.mark.needs_reference


def _cfg(d):
    return ref_shims.AttrDict(d)


@pytest.mark.parametrize("dataset,target", [
    ("ns", "dyffusion_b200.backbones.unet_simple.UNet"),
    ("sst", "dyffusion_b200.backbones.unet.Unet"),
    ("spring", "dyffusion_b200.backbones.simple_conv_net.SimpleConvNet"),
])
def test_reference_experiments_accept_the_dropins(dataset, target):
    ref_shims.install()
    from src.experiment_types.forecasting_multi_horizon import MultiHorizonForecastingDYffusion
    from src.experiment_types.interpolation import InterpolationExperiment
    from src.utilities.naming import clean_name

    horizon = 4
    mk = dict(C.MODELS[dataset]["kwargs"])
    mc = lambda: _cfg(dict(_target_=target, name="", verbose=False, loss_function="mse", **mk))
    dm = _cfg(dict(C.DATASETS[dataset]["datamodule"], horizon=horizon))
    ipol = InterpolationExperiment(model_config=mc(), datamodule_config=dm, verbose=False, num_predictions=1)
    dk = C.diffusion_kwargs(dataset, horizon=horizon)
    if dataset == "sst":
        dk["additional_interpolation_steps"] = 2
    dc = _cfg(dict(_target_="dyffusion_b200.diffusion.dyffusion.DYffusion", **dk))
    dc["interpolator"] = ipol
    exp = MultiHorizonForecastingDYffusion(model_config=mc(), datamodule_config=dm, diffusion_config=dc, verbose=False,
                                           num_predictions=1)
    diff = exp.model
    assert type(diff).__module__ == "dyffusion_b200.diffusion.dyffusion"
    assert clean_name(target) in ("SimpleUnet", "UNetR", "SimpleCNN")
    cin, ccond, cout = C.channels(dataset, "F", dk["forward_conditioning"])
    assert (diff.model.num_input_channels, diff.model.num_conditional_channels, diff.model.num_output_channels) == \
        (cin, ccond, cout)
    icin, iccond, _ = C.channels(dataset, "I", dk["forward_conditioning"])
    assert (ipol.model.num_input_channels, ipol.model.num_conditional_channels) == (icin, iccond)
    assert diff.interpolator is ipol and all(not p.requires_grad for p in ipol.parameters())
    assert diff.num_timesteps == horizon + dk["additional_interpolation_steps"]
    assert diff.num_params == diff.model.num_params > 0         # frozen interpolator: only the forecaster trains
    # keys the reference's checkpoint code relies on (interface.py:160-161, forecasting_multi_horizon.py:422-424)
    keys = list(exp.state_dict().keys())
    assert any(k.startswith("model.model.") for k in keys) and any(k.startswith("model.interpolator.model.") for k in keys)
    with pytest.raises(Exception):  # no GPU in the build container: compute must fail loudly, never fall back
        exp.predict(torch.zeros(1, C.DATASETS[dataset]["channels"], *C.DATASETS[dataset]["spatial"]),
                    condition=None if not C.DATASETS[dataset]["static"] else
                    torch.zeros(1, C.DATASETS[dataset]["static"], *C.DATASETS[dataset]["spatial"]))
