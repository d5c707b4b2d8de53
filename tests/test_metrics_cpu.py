"""CPU tests of the ensemble-metrics oracle (oracle/metrics_oracle.py): the sorted-ensemble CRPS algorithm the reference
reaches through xskillscore/properscoring against the closed form it integrates, hand-computed cases, and the
evaluation.py bookkeeping (shapes, keys, mean_over_samples)."""
import numpy as np

from oracle import metrics_oracle as M


def test_crps_sorted_algorithm_equals_energy_form():
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 50):
        fc = rng.normal(size=(7, 11, n))
        obs = rng.normal(size=(7, 11))
        np.testing.assert_allclose(M.crps_ensemble(obs, fc), M.crps_energy_form(obs, fc), rtol=1e-12, atol=1e-12)
    fc = np.round(rng.normal(size=(40, 6)), 1)  # ties, and observations equal to members
    obs = fc[:, 2].copy()
    np.testing.assert_allclose(M.crps_ensemble(obs, fc), M.crps_energy_form(obs, fc), rtol=1e-12, atol=1e-12)


def test_crps_known_answers():
    # one member: CRPS = |x - y|;  two members {0, 1}, y = 0.5: E|X-y| = 0.5, E|X-X'| = 0.5 -> 0.25
    assert M.crps_ensemble(np.array(2.0), np.array([0.5])) == 1.5
    assert abs(M.crps_ensemble(np.array(0.5), np.array([0.0, 1.0])) - 0.25) < 1e-15
    # observation below / above the whole ensemble
    assert abs(M.crps_ensemble(np.array(-1.0), np.array([0.0, 1.0])) - (1.5 - 0.25)) < 1e-15
    assert abs(M.crps_ensemble(np.array(3.0), np.array([0.0, 1.0])) - (2.5 - 0.25)) < 1e-15


def test_evaluate_ensemble_prediction_bookkeeping():
    rng = np.random.default_rng(1)
    preds = rng.normal(size=(5, 4, 3, 6, 7)).astype(np.float32)
    tg = rng.normal(size=(4, 3, 6, 7)).astype(np.float32)
    r = M.evaluate_ensemble_prediction(preds, tg, also_per_member_metrics=True)
    assert set(r) == {"ssr", "crps", "mse", "mse_per_mem", "mse_per_mem_mean"}
    assert np.isclose(r["mse"], ((preds.mean(0) - tg) ** 2).mean())
    assert r["mse_per_mem"].shape == (5,) and np.isclose(r["mse_per_mem_mean"], ((preds - tg) ** 2).mean())
    assert np.isclose(r["ssr"], np.sqrt(preds.astype(np.float64).var(0).mean()) / np.sqrt(r["mse"]))
    per = M.evaluate_ensemble_prediction(preds, tg, mean_over_samples=False)
    assert per["crps"].shape == per["mse"].shape == per["ssr"].shape == (4,)
    assert np.isclose(per["crps"].mean(), r["crps"]) and np.isclose(per["mse"].mean(), r["mse"])
    # a 2-D target gets a channel axis (evaluation.py:35-38)
    r2 = M.evaluate_ensemble_prediction(preds[:, :, 0, 0, :], tg[:, 0, 0, :])
    assert np.isfinite(r2["crps"])


# ---------------------------------------------------------------- pinned against the reference's own evaluation.py
import importlib.util  # noqa: E402
import sys  # noqa: E402
import types  # noqa: E402

import pytest  # noqa: E402


def _reference_evaluation():
    """`src/utilities/evaluation.py` of the reference, imported UNMODIFIED with two stub modules in place of the absent
    third-party packages: `xarray.DataArray` only carries (values, dims); `xskillscore.crps_ensemble` is the published
    properscoring algorithm as restated in oracle/metrics_oracle.py (members moved last, mean over `dim`).  Everything else --
    ensemble-mean MSE, per-member MSE, spread-skill ratio, channel-axis / mean_over_samples bookkeeping -- is the reference's
    own code."""
    xr = types.ModuleType("xarray")

    class DataArray:
        def __init__(self, values, dims):
            self.values, self.dims = np.asarray(values), list(dims)

    xr.DataArray = DataArray
    xs = types.ModuleType("xskillscore")

    def crps_ensemble(observations, forecasts, member_dim="member", dim=None):
        fc = np.moveaxis(forecasts.values, forecasts.dims.index(member_dim), -1)
        out = M.crps_ensemble(observations.values, fc)
        axes = tuple(observations.dims.index(d) for d in (dim or []))
        return DataArray(out.mean(axis=axes) if axes else out, [d for d in observations.dims if d not in (dim or [])])

    xs.crps_ensemble = crps_ensemble
    saved = {k: sys.modules.get(k) for k in ("xarray", "xskillscore")}
    sys.modules["xarray"], sys.modules["xskillscore"] = xr, xs
    try:
        spec = importlib.util.spec_from_file_location("_ref_evaluation", "/root/reference/src/utilities/evaluation.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


@pytest.mark.needs_reference
@pytest.mark.parametrize("shape", [(5, 4, 3, 6, 7), (3, 6, 12), (8, 2, 1, 10, 10)])
@pytest.mark.parametrize("mean_over_samples", [True, False])
def test_oracle_equals_the_reference_evaluation_module(shape, mean_over_samples):
    ref = _reference_evaluation()
    rng = np.random.default_rng(7)
    preds = rng.normal(size=shape)
    tg = rng.normal(size=shape[1:])
    want = ref.evaluate_ensemble_prediction(preds, tg, also_per_member_metrics=True, mean_over_samples=mean_over_samples)
    got = M.evaluate_ensemble_prediction(preds, tg, also_per_member_metrics=True, mean_over_samples=mean_over_samples)
    assert set(got) == set(want)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-12, atol=1e-14, err_msg=k)
    if mean_over_samples:  # the stand-alone SSR entry point (evaluation.py:100-120) with its own RMSE
        np.testing.assert_allclose(ref.evaluate_ensemble_spread_skill_ratio(preds, tg, mean_dims=None), got["ssr"], rtol=1e-12)
