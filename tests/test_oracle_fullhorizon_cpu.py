"""Pins the oracle at the BASELINE.json horizons: oracle/dyffusion_oracle.py against goldens the UNMODIFIED reference
produced for the same synthetic weights / inputs (tests/golden/make_golden_full.py; stored as fp16, hence 1e-3)."""
import os

import pytest
import torch

from oracle import configs as C
from oracle import dyffusion_oracle as O
from oracle.synth import synth_state_dict, synth_tensor
from tests import helpers as H

KAT = H.golden_json("schedule_kat_full.json")
SHAPES = H.golden_json("state_shapes.json")


@pytest.mark.parametrize("name", sorted(KAT))
def test_oracle_matches_reference_golden_at_full_horizon(name):
    meta = KAT[name]
    ds = meta["dataset"]
    g = H.golden_pt(f"sample_{name}.pt")
    dk = C.diffusion_kwargs(ds, **meta["overrides"])
    nets = []
    for role, seed in (("F", 3), ("I", 2)):
        sd = synth_state_dict(SHAPES[f"{ds}_{role}"], seed=seed)
        for k, f in meta["weight_scale"].items():
            sd[k] = sd[k] * f
        nets.append(H.oracle_net(ds, role, sd))
    ic, static = H.sampler_case_inputs(name, ds, g["rows"])
    n = {"i": 0}

    def noise(t):
        n["i"] += 1
        return synth_tensor(f"{name}.noise{n['i'] - 1}", tuple(t.shape))

    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        out = O.sample_loop(nets[0], nets[1], H.oracle_schedule(dk), ic, static,
                            num_input_channels=C.DATASETS[ds]["channels"], forward_conditioning=dk["forward_conditioning"],
                            refine_intermediate_predictions=dk["refine_intermediate_predictions"], noise_fn=noise)
    assert sorted(out) == meta["keys"] and n["i"] == meta["noise_draws"]
    sched = H.oracle_schedule(dk)
    assert [float(s) for s in sched.sampling_schedule] == meta["sampling_schedule"]
    for k, v in g["preds"].items():
        assert H.rel_l2(out[k], v.float()) <= 1e-3, (k, H.rel_l2(out[k], v.float()))
