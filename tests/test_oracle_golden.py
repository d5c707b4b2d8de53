"""CPU: pins oracle/ (the torch-fp32 restatement) against the golden vectors produced by the real reference
(tests/golden/make_golden.py).  Tolerance: fp32 vs fp32 of the same math -> rel-L2 <= 2e-5 per forward and
<= 2e-4 per chained trajectory (summation-order differences only)."""
import pytest
import torch

from oracle import configs as C
from oracle import dyffusion_oracle as O
from oracle.synth import SiteDropout, synth_state_dict, synth_tensor
from tests import helpers as H

SHAPES = H.golden_json("state_shapes.json")
KAT = H.golden_json("schedule_kat.json")


@pytest.mark.parametrize("dataset", ["ns", "sst", "spring"])
@pytest.mark.parametrize("role", ["F", "I"])
def test_forward_matches_reference_golden(dataset, role):
    tag = f"{dataset}_{role}"
    g = H.golden_pt(f"fwd_{tag}.pt")
    sd = synth_state_dict(SHAPES[tag], seed=g["weight_seed"])
    x, cond = H.forward_inputs(dataset, role, rows=g["rows"])
    with torch.no_grad():
        y = H.oracle_net(dataset, role, sd)(x, g["time"], cond)
        y_drop = H.oracle_net(dataset, role, sd, drop=SiteDropout(g["drop_seed"]))(x, g["time"], cond)
    assert H.rel_l2(y, g["y"]) <= 2e-5
    assert H.rel_l2(y_drop, g["y_sitedrop"]) <= 2e-5  # pins dropout placement + 1/(1-p) scaling
    assert H.rel_l2(g["y"], g["y_sitedrop"]) > 1e-2  # the dropout golden is not vacuous


SAMPLER_CASES = [k for k in KAT if k != "schedules"]


@pytest.mark.parametrize("name", SAMPLER_CASES)
def test_sampler_matches_reference_golden(name):
    meta = KAT[name]
    ds = meta["dataset"]
    dk = C.diffusion_kwargs(ds, **meta["overrides"])
    g = H.golden_pt(f"sample_{name}.pt")
    sched = H.oracle_schedule(dk)
    assert [float(s) for s in sched.sampling_schedule] == meta["sampling_schedule"]
    assert sched.num_timesteps == meta["num_timesteps"]
    assert O.count_calls(sched, dk["refine_intermediate_predictions"], dk["sampling_type"],
                         dk["use_cold_sampling_for_last_step"]) == (meta["calls_F"], meta["calls_I"])
    sdI = synth_state_dict(SHAPES[f"{ds}_I"], seed=2)
    sdF = synth_state_dict(SHAPES[f"{ds}_F"], seed=3)
    ic, static = H.sampler_case_inputs(name, ds, g["rows"])
    n = {"i": 0}

    def noise(t):
        n["i"] += 1
        return synth_tensor(f"{name}.noise{n['i'] - 1}", tuple(t.shape))

    with torch.no_grad():
        out = O.sample_loop(H.oracle_net(ds, "F", sdF), H.oracle_net(ds, "I", sdI), sched, ic, static,
                            num_input_channels=C.DATASETS[ds]["channels"],
                            forward_conditioning=dk["forward_conditioning"], sampling_type=dk["sampling_type"],
                            time_encoding=dk["time_encoding"],
                            refine_intermediate_predictions=dk["refine_intermediate_predictions"],
                            use_cold_sampling_for_last_step=dk["use_cold_sampling_for_last_step"], noise_fn=noise)
    assert sorted(out) == meta["keys"]
    assert n["i"] == meta["noise_draws"]
    for k, v in g["preds"].items():
        assert H.rel_l2(out[k], v) <= 2e-4, k


def test_schedule_known_answers():
    """Host logic must be EXACT: step->time map, schedule strings, error behaviour."""
    n_ok = n_err = 0
    for rec in KAT["schedules"]:
        a = rec["args"]
        kw = dict(timesteps=a["timesteps"], schedule=a.get("schedule", "before_t1_only"),
                  additional_interpolation_steps=a.get("additional_interpolation_steps", 0),
                  additional_interpolation_steps_factor=a.get("additional_interpolation_steps_factor", 0),
                  interpolate_before_t1=a.get("interpolate_before_t1", True), sampling_schedule=rec["spec"])
        if "error" in rec:
            with pytest.raises((AssertionError, ValueError, IndexError)):
                O.Schedule(**kw)
            n_err += 1
            continue
        s = O.Schedule(**kw)
        assert s.num_timesteps == rec["num_timesteps"]
        assert [float(v) for v in s.sampling_schedule] == rec["schedule"], (a, rec["spec"])
        assert all(isinstance(v, int) for v in s.sampling_schedule) == rec["all_int"]
        assert [float(s.tau(d)) for d in range(s.num_timesteps)] == rec["tau"]
        assert {str(k): float(v) for k, v in s.dynamical_steps.items()} == rec["dynamical"]
        n_ok += 1
    assert n_ok + n_err == len(KAT["schedules"]) and n_ok >= 30
    with pytest.raises(ValueError):  # dyffusion.py:297-298
        O.Schedule(7, additional_interpolation_steps=25, sampling_schedule="bogus")
    with pytest.raises(AssertionError):  # dyffusion.py:46-47 horizon must be > 1
        O.Schedule(1)


def test_call_counts_at_baseline_configs():
    """NS h=16 refine -> 16 F + 44 I; SST h=7,k=25 -> 32 F + 61 I; spring h=134 refine -> 134 F + 398 I."""
    for ds, want in (("ns", (16, 44)), ("sst", (32, 61)), ("spring", (134, 398))):
        dk = C.diffusion_kwargs(ds)
        assert O.count_calls(H.oracle_schedule(dk), dk["refine_intermediate_predictions"]) == want
