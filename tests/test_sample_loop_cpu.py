"""CPU pinning of the Python-driven sampling loop of the drop-in (`DYffusion._sample_loop_python`, the route for
`log_every_t`, schedules that stop early and foreign interpolators) and of `sample_loop`'s bookkeeping (SURVEY.md Appendix E:
output keys, return triple, refinement with fractional `prediction_timesteps`, window > 1) against the reference's own
`BaseDYffusion.sample_loop` (src/diffusion/dyffusion.py:335-431).

The drop-in is wrapped around the REFERENCE's torch backbones (its host logic only calls `predict_forward` / `predict`), so
with equal seeds every returned tensor must equal the reference's bit for bit, interpolator dropout on."""
import pytest
import torch

from oracle import configs as C
from tests import helpers as H

pytestmark = pytest.mark.needs_reference


def _pair(dataset, horizon, window=1, **ov):
    from oracle import ref_build, ref_shims
    from tests.golden.make_golden import load_synth
    from dyffusion_b200.diffusion import DYffusion
    saved = C.DATASETS[dataset]["datamodule"]["window"]
    C.DATASETS[dataset]["datamodule"]["window"] = window  # the reference derives the interpolator's channels from it
    try:
        ipol = ref_build.build_interpolator(dataset, horizon=horizon)
        exp = ref_build.build_dyffusion(dataset, ipol, horizon=horizon, **ov)
    finally:
        C.DATASETS[dataset]["datamodule"]["window"] = saved
    load_synth(ipol.model, seed=2), load_synth(exp.model.model, seed=3)
    dk = C.diffusion_kwargs(dataset, horizon=horizon, **ov)
    mine = DYffusion(model=exp.model.model, interpolator=ipol, verbose=False, **dk).eval()
    return exp.model.eval(), mine, ipol


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    return torch.equal(a, b)


CASES = {
    "full_refine": dict(horizon=4),
    "log_every_t": dict(horizon=4, log_every_t=1),
    "log_auto_naive": dict(horizon=3, log_every_t="auto", sampling_type="naive", refine_intermediate_predictions=False),
    "stops_early": dict(horizon=5, sampling_schedule=[0, 1, 2], refine_intermediate_predictions=False),
    "aux_steps_every2nd": dict(horizon=3, additional_interpolation_steps=4, sampling_schedule="every2nd"),
    "cold_last_step": dict(horizon=4, use_cold_sampling_for_last_step=True, log_every_t=1),
    "fractional_refine": dict(horizon=3, prediction_timesteps=[0.5, 1, 1.5, 2, 2.5]),
}


@pytest.mark.parametrize("name", list(CASES))
def test_sample_loop_equals_reference(name):
    ov = dict(CASES[name])
    ref, mine, _ = _pair("spring", ov.pop("horizon"), **ov)
    ic, static = H.sampler_case_inputs(f"loop.{name}", "spring", 3)
    outs = []
    for d in (ref, mine):
        torch.manual_seed(13)
        with torch.no_grad():
            outs.append(d.sample_loop(ic, static_condition=static))
    (a0, ad, a2), (b0, bd, b2) = outs
    assert list(ad) == list(bd), (list(ad), list(bd))
    for k in ad:
        assert _same(ad[k], bd[k]), k
    assert _same(a0, b0) and _same(a2, b2)
    torch.manual_seed(13)
    with torch.no_grad():
        s = mine.sample(ic, static_condition=static)
    assert list(s) == list(ad) and all(_same(s[k], ad[k]) for k in ad)


def test_window_two_starts_from_the_last_frame():
    """`x_s` starts from the LAST C channels of a window stacked in channels (dyffusion.py:348)."""
    ref, mine, _ = _pair("spring", 3, window=2)
    ic = H.synth_tensor("loop.w2.ic", (2, 8, 10, 10))
    static = H.synth_tensor("loop.w2.static", (2, 1, 10, 10), kind="mask")
    outs = []
    for d in (ref, mine):
        torch.manual_seed(3)
        with torch.no_grad():
            outs.append(d.sample(ic, static_condition=static))
    assert list(outs[0]) == list(outs[1]) == ["t1_preds", "t2_preds", "t3_preds"]
    assert all(torch.equal(outs[0][k], outs[1][k]) and outs[1][k].shape == (2, 4, 10, 10) for k in outs[0])


def test_interpolation_time_is_checked_like_the_reference():
    ref, mine, _ = _pair("spring", 3)
    ic, static = H.sampler_case_inputs("loop.err", "spring", 2)
    for d in (ref, mine):
        with pytest.raises(AssertionError):
            d.q_sample(x0=ic, x_end=ic, t=None, interpolation_time=torch.full((2,), 3.0), static_condition=static)
        with pytest.raises(AssertionError):
            d.q_sample(x0=ic, x_end=ic, t=torch.ones(2), interpolation_time=torch.ones(2), static_condition=static)
        with pytest.raises(AssertionError):
            d.predict_x_last(condition=ic, x_t=ic, t=torch.full((2,), 3.0), static_condition=static)
