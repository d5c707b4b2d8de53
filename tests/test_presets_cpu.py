"""The product's shipped configurations (dyffusion_b200/configs/*.yaml, read by dyffusion_b200.presets) against the test
infrastructure's statement of the reference's settings (oracle/configs.py) and, where /root/reference exists, against the
reference's own YAML files key by key."""
import os

import pytest
import yaml

from dyffusion_b200 import presets as P
from oracle import configs as C

REF = "/root/reference/src/configs"


@pytest.mark.parametrize("name", ["ns", "sst", "spring"])
def test_preset_equals_oracle_configs(name):
    p = P.load_preset(name)
    want_model = C.MODELS[name]["kwargs"]
    for k, v in want_model.items():
        assert p["model"][k] == v, (k, p["model"].get(k), v)
    for k, v in dict(want_model, **C.INTERPOLATOR_OVERRIDES[name]).items():
        assert p["interpolator_model"][k] == v, k
    dk = C.diffusion_kwargs(name)
    for k, v in dk.items():
        assert p["diffusion"][k] == v, (k, p["diffusion"].get(k), v)
    fc = dk["forward_conditioning"]
    for role in ("F", "I"):
        assert P.channels(p, role) == C.channels(name, role, fc)
    d = C.DATASETS[name]
    assert tuple(p["dataset"]["spatial_shape"]) == tuple(d["spatial"]) and p["dataset"]["channels"] == d["channels"]
    assert p["model_target"].rsplit(".", 1)[1] == C.MODELS[name]["ref_target"].rsplit(".", 1)[1]


@pytest.mark.parametrize("name,steps,keys", [("ns", 16, 16), ("sst", 32, 7), ("spring", 134, 134)])
def test_preset_builds_the_drop_in_on_the_host(name, steps, keys):
    dyf = P.build_dyffusion(name, device="cpu")
    assert len(dyf.sampling_schedule) == steps and len(dyf.dynamical_steps) + 1 >= keys
    assert "dyffusion" in type(dyf).__module__.lower()  # the reference's channel bookkeeping keys on it
    sd = dyf.model.state_dict()
    assert any(k.endswith("running_var") and float(v.min()) >= 0.5 and float(v.std()) > 0 for k, v in sd.items()) or name == "sst"


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
@pytest.mark.parametrize("ours,theirs", [("model/unet_simple_navier_stokes_b200.yaml", "model/unet_simple_navier_stokes.yaml"),
                                         ("model/unet_resnet_b200.yaml", "model/unet_resnet.yaml"),
                                         ("model/cnn_simple_b200.yaml", "model/cnn_simple.yaml"),
                                         ("diffusion/dyffusion_b200.yaml", "diffusion/dyffusion.yaml")])
def test_hydra_files_differ_from_the_reference_only_in_the_target(ours, theirs):
    a = yaml.safe_load(open(os.path.join(P.CONFIG_DIR, ours)))
    b = yaml.safe_load(open(os.path.join(REF, theirs)))
    sect = "diffusion" if "diffusion" in b else "model" if "model" in b else None
    a, b = (a[sect], b[sect]) if sect else (a, b)
    extra = {"max_rows_per_call", "cuda_graph"}
    assert set(a) - extra == set(b)
    for k in b:
        if k == "_target_":
            assert a[k].startswith("dyffusion_b200.") and a[k].rsplit(".", 1)[1] == b[k].rsplit(".", 1)[1]
        elif k != "defaults":
            assert a[k] == b[k], k
