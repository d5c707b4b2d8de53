"""CPU tests of the autoregressive rollout (second half).

* the oracle restatement (oracle/rollout_oracle.py over the oracle sampler and the oracle boundary conditions) against the
  committed reference goldens (tests/golden/rollout_*.pt, made by tests/golden/make_rollout_golden.py);
* the product's host logic (`dyffusion_b200.rollout.MultiHorizonRollout`) driven around a CPU sampler -- it only touches
  the `predict_forward / sample_loop` surface -- against the oracle, and, where /root/reference is importable, bit for bit
  against the reference's own `_evaluation_step` with interpolator dropout ON (same torch RNG stream)."""
import functools
import inspect

import numpy as np
import pytest
import torch

from oracle import boundary_oracle, rollout_oracle
from tests import helpers as H

from dyffusion_b200.rollout import MultiHorizonRollout


class _OracleDiffusion:
    """The oracle sampler behind the two methods the rollout calls on a diffusion object."""

    def __init__(self, dataset, horizon):
        self._sample = H.oracle_rollout_sampler(dataset, horizon)
        self.num_timesteps = horizon
        self.hparams = {"timesteps": horizon}
        self.calls = []

    def sample_loop(self, initial_condition, static_condition=None, log_every_t=None, num_predictions=None):
        raise NotImplementedError

    def predict_forward(self, inputs, condition=None, metadata=None, **kwargs):
        self.calls.append((tuple(inputs.shape), kwargs.get("num_predictions")))
        return self._sample(inputs, condition)


def _oracle_run(name):
    c, batch, t0, dt = H.rollout_case(name)
    bc = functools.partial(boundary_oracle.boundary_conditions, c["system"])
    out = rollout_oracle.evaluation_step(
        H.oracle_rollout_sampler(c["dataset"], c["horizon"]), batch, horizon=c["horizon"], num_predictions=c["members"],
        autoregressive_steps=c["ar_steps"], boundary_conditions=lambda preds, targets, metadata, time: bc(
            preds, targets, metadata, time=time), t0=t0, dt=dt)
    return c, batch, t0, dt, out


@pytest.mark.parametrize("name", list(H.ROLLOUT_CASES))
def test_oracle_rollout_matches_reference_golden(name):
    c, batch, _, _, out = _oracle_run(name)
    gold = H.golden_pt(f"{name}.pt")["preds"]
    assert sorted(k for k in out if k.endswith("_preds")) == sorted(gold)
    T = c["horizon"] * (c["ar_steps"] + 1)
    for t in range(1, T + 1):
        got, want = torch.from_numpy(out[f"t{t}_preds"]), gold[f"t{t}_preds"]
        assert got.shape == want.shape
        assert H.rel_l2(got, want) <= 1e-3, (name, t, H.rel_l2(got, want))  # chained fp32 trajectories (<= 3 sampler calls)
        assert np.array_equal(out[f"t{t}_targets"], batch["dynamics"][:, t].numpy())
        if c["system"] == "spring-mesh":  # the masked writes themselves are exact
            fm = batch["metadata"]["fixed_mask"]
            assert torch.equal(got[..., fm], want[..., fm]) if got.ndim == 4 else torch.equal(got[:, fm], want[:, fm])


@pytest.mark.parametrize("name", list(H.ROLLOUT_CASES))
def test_product_host_logic_equals_oracle(name):
    c, batch, t0, dt, want = _oracle_run(name)
    diff = _OracleDiffusion(c["dataset"], c["horizon"])
    ro = MultiHorizonRollout(diff, horizon=c["horizon"], num_predictions=c["members"], autoregressive_steps=c["ar_steps"])
    bc = functools.partial(boundary_oracle.boundary_conditions, c["system"])
    before = batch["dynamics"].clone()
    got = ro.evaluation_step(batch, "test", boundary_conditions=bc, t0=t0, dt=dt)
    assert list(got) == list(want)
    for k in want:
        assert torch.equal(got[k], torch.from_numpy(want[k])), k
    assert torch.equal(batch["dynamics"], before)  # documented difference: the caller's batch is left alone
    rows = c["members"] * c["batch"]
    assert [r[0][0] for r in diff.calls] == [rows] * (c["ar_steps"] + 1)  # ONE sampler call per autoregressive step
    assert [r[1] for r in diff.calls] == [c["members"]] + [1] * c["ar_steps"]  # :160 `num_predictions=1` on AR steps
    # numpy mode and the stacking `test_step` does (forecasting_multi_horizon.py:246-249)
    got_np = ro.evaluation_step(batch, "test", boundary_conditions=bc, t0=t0, dt=dt, to_numpy=True)
    assert all(isinstance(v, np.ndarray) for v in got_np.values())
    p, t = ro.stack_trajectory(got)
    T = c["horizon"] * (c["ar_steps"] + 1)
    assert tuple(t.shape) == (T, c["batch"], *batch["dynamics"].shape[2:])
    assert tuple(p.shape) == ((c["members"],) if c["members"] > 1 else ()) + tuple(t.shape)
    pn, tn = ro.stack_trajectory(got_np)
    assert np.array_equal(pn, p.numpy()) and np.array_equal(tn, t.numpy())


def test_member_major_rows_and_single_pass():
    c, batch, t0, dt = H.rollout_case("rollout_spring")
    diff = _OracleDiffusion("spring", c["horizon"])
    ro = MultiHorizonRollout(diff, horizon=c["horizon"], num_predictions=3, autoregressive_steps=2)
    x = ro.transform_inputs(ro.get_inputs_from_dynamics(batch["dynamics"]), split="test", ensemble=True)
    assert x.shape[0] == 6 and torch.equal(x[0:2], x[2:4]) and torch.equal(x[0:2], batch["dynamics"][:, 0])
    out = ro.evaluation_step(batch, "val", autoregressive=False)  # first validation loader: one pass (:135-136)
    assert sorted(out, key=lambda k: (int(k[1:].split("_")[0]), k)) == [f"t{t}_{s}" for t in range(1, 5)
                                                                           for s in ("preds", "targets")]
    assert ro.num_autoregressive_steps == 2 and ro.prediction_horizon == 12


def test_errors_follow_the_reference():
    c, batch, t0, dt = H.rollout_case("rollout_spring_single")
    diff = _OracleDiffusion("spring", c["horizon"])
    with pytest.raises(AssertionError):
        MultiHorizonRollout(diff, horizon=3, autoregressive_steps=-1)
    with pytest.raises(AssertionError):
        MultiHorizonRollout(diff, horizon=3, autoregressive_steps=1, prediction_horizon=6)
    with pytest.raises(AssertionError):
        MultiHorizonRollout(diff, horizon=4)  # "diffusion timesteps must be equal to horizon"
    ro = MultiHorizonRollout(diff, horizon=3, prediction_horizon=30)  # longer than the batch holds
    assert ro.num_autoregressive_steps == 9
    with pytest.raises(ValueError):
        ro.evaluation_step(batch, "test")
    ro = MultiHorizonRollout(diff, horizon=3, prediction_horizon=5)  # not a multiple of the horizon: stops after t5 (:153-156)
    out = ro.evaluation_step(batch, "test")
    assert [k for k in out if k.endswith("preds")] == [f"t{t}_preds" for t in range(1, 6)]
    with pytest.raises(AssertionError):
        ro.prediction_timesteps = [1, 2, 7]
    with pytest.raises(AssertionError):
        ro.get_preds_at_t_for_batch(batch, 4, "test")


@pytest.mark.needs_reference
@pytest.mark.parametrize("name", ["rollout_spring", "rollout_spring_single"])
def test_product_and_oracle_equal_the_reference_loop_with_dropout(name):
    """Bit for bit against the reference's own loop, interpolator dropout ON: the product's host logic drives the
    REFERENCE DYffusion module here, so equal RNG consumption order is part of what is checked."""
    from tests.golden.make_rollout_golden import reference_rollout
    want, exp, bc, batch, t0, dt = reference_rollout(name, dropout=True, seed=11)
    c = H.ROLLOUT_CASES[name]
    ro = MultiHorizonRollout(exp.model, horizon=c["horizon"], num_predictions=c["members"],
                             autoregressive_steps=c["ar_steps"])
    torch.manual_seed(11)
    got = ro.evaluation_step(batch, "test", boundary_conditions=bc, t0=t0, dt=dt, to_numpy=True)
    assert list(got) == list(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    torch.manual_seed(11)
    sample = lambda ic, static: exp.model.predict_forward(ic, condition=static)
    orc = rollout_oracle.evaluation_step(sample, batch, horizon=c["horizon"], num_predictions=c["members"],
                                         autoregressive_steps=c["ar_steps"], boundary_conditions=bc, t0=t0, dt=dt)
    for k in want:
        assert np.array_equal(orc[k], want[k]), k


@pytest.mark.needs_reference
def test_fractional_prediction_timesteps_equal_the_reference():
    """`prediction_timesteps` with non-integer steps: refinement emits `t0.5_preds`-style keys, targets are None there
    (forecasting_multi_horizon.py:166-170, dyffusion.py:408-422)."""
    from oracle import ref_build
    from tests.golden.make_golden import load_synth
    c, batch, _, _ = H.rollout_case("rollout_spring_single")
    ipol = ref_build.build_interpolator("spring", horizon=3)
    exp = ref_build.build_dyffusion("spring", ipol, horizon=3)
    load_synth(ipol.model, seed=2), load_synth(exp.model.model, seed=3)
    steps = [0.5, 1, 1.5, 2, 3]
    exp._prediction_timesteps = list(steps)
    exp.hparams.autoregressive_steps = 1
    torch.manual_seed(5)
    want = exp._evaluation_step({k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}, 0, "test")
    ro = MultiHorizonRollout(exp.model, horizon=3, autoregressive_steps=1, prediction_timesteps=steps)
    torch.manual_seed(5)
    got = ro.evaluation_step(batch, "test", to_numpy=True)
    assert list(got) == list(want) and "t0.5_preds" in got and got["t0.5_targets"] is None
    for k in want:
        assert (got[k] is None and want[k] is None) or np.array_equal(got[k], want[k]), k
