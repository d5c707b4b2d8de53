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
