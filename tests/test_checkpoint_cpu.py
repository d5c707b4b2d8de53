"""CPU tests of the offline checkpoint ingest (dyffusion_b200/checkpoint.py): key splitting / renaming on
synthetic Lightning-style state dicts built from the committed reference key lists (tests/golden/state_shapes.json), and --
where /root/reference is importable -- against the state dict of the REAL reference Lightning modules."""
import pytest
import torch

from dyffusion_b200.checkpoint import load_reference_checkpoint, rename_state_dict_keys, split_state_dict
from oracle import ref_shims
from tests import helpers as H

SHAPES = H.golden_json("state_shapes.json")


def _fake(tag, prefix):
    return {prefix + k: torch.full(tuple(s) if s else (), float(i % 7)) for i, (k, s) in enumerate(SHAPES[tag].items())}


def test_split_dyffusion_and_plain_runs():
    sd = {**_fake("ns_F", "model.model."), **_fake("ns_I", "model.interpolator.model."), "model_ema.shadow.0": torch.zeros(1)}
    parts = split_state_dict(sd)
    assert sorted(parts["model"]) == sorted(SHAPES["ns_F"]) and sorted(parts["interpolator"]) == sorted(SHAPES["ns_I"])
    plain = split_state_dict(_fake("sst_I", "model."))
    assert sorted(plain["model"]) == sorted(SHAPES["sst_I"]) and "interpolator" not in plain
    with pytest.raises(ValueError):
        split_state_dict({"optimizer.foo": torch.zeros(1)})


def test_legacy_qkv_rename():
    old = {k.replace("fn.to_qkv.1.weight", "fn.to_qkv.weight"): v for k, v in _fake("sst_F", "model.model.").items()}
    assert any("downs.0.2.fn.fn.to_qkv.weight" in k for k in old)
    sd, renamed = rename_state_dict_keys(dict(old))
    assert renamed and "model.model.mid_attn.fn.fn.to_qkv.weight" in sd  # the bottleneck attention keeps its name
    assert sorted(split_state_dict(old)["model"]) == sorted(SHAPES["sst_F"])


def test_load_into_dropin_backbones(tmp_path):
    from tests.test_host_cpu import _build  # engine-backed drop-in classes on CPU (parameters only, no compute)
    F, I = _build("spring", "F"), _build("spring", "I")
    ckpt = {"state_dict": {**{"model.model." + k: torch.randn_like(v) if v.is_floating_point() else v for k, v in F.state_dict().items()},
                           **{"model.interpolator.model." + k: torch.randn_like(v) if v.is_floating_point() else v
                              for k, v in I.state_dict().items()}}, "epoch": 7, "global_step": 123}
    path = str(tmp_path / "last.ckpt")
    torch.save(ckpt, path)
    out = load_reference_checkpoint(path, model=F, interpolator=I)
    assert out["epoch"] == 7 and out["global_step"] == 123
    for k, v in F.state_dict().items():
        assert torch.equal(v, ckpt["state_dict"]["model.model." + k]), k
    for k, v in I.state_dict().items():
        assert torch.equal(v, ckpt["state_dict"]["model.interpolator.model." + k]), k


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("dataset", ["spring", "sst"])
def test_against_the_real_reference_modules(dataset):
    """The prefixes are not guessed: build the reference's InterpolationExperiment + MultiHorizonForecastingDYffusion and
    split THEIR state dict."""
    from oracle import ref_build
    interp = ref_build.build_interpolator(dataset, horizon=4 if dataset == "spring" else 3)
    exp = ref_build.build_dyffusion(dataset, interp, horizon=4) if dataset == "spring" else ref_build.build_dyffusion(dataset, interp, horizon=3, **{})
    parts = split_state_dict(exp.state_dict())
    assert sorted(parts["model"]) == sorted(exp.model.model.state_dict())
    assert sorted(parts["interpolator"]) == sorted(interp.model.state_dict())
    for k, v in exp.model.model.state_dict().items():
        assert torch.equal(parts["model"][k], v)
    plain = split_state_dict(interp.state_dict())
    assert sorted(plain["model"]) == sorted(interp.model.state_dict())


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("kind", ["dyffusion", "plain"])
def test_ema_weights_against_the_real_litema(kind):
    """`use_ema=True`: the shadow names are not guessed either -- the reference's own `LitEma` (src/models/modules/ema.py) is
    built over the reference module, stepped, and its buffers are what a `use_ema` run stores under `model_ema.`; the weights
    the reference would evaluate with (`LitEma.copy_to`, via `ema_scope`) must be what the ingest returns."""
    from oracle import ref_build
    ref_shims.install()
    from src.models.modules.ema import LitEma
    interp = ref_build.build_interpolator("spring", horizon=3)
    exp = ref_build.build_dyffusion("spring", interp, horizon=3) if kind == "dyffusion" else interp
    ema = LitEma(exp.model, decay=0.9)
    with torch.no_grad():
        for p in exp.model.parameters():
            if p.requires_grad:
                p.add_(0.5 * torch.randn_like(p))
    ema(exp.model)  # one EMA update: the shadows now differ from both the initial and the current weights
    sd = {k: v.clone() for k, v in {**exp.state_dict(), **{"model_ema." + k: v for k, v in ema.state_dict().items()}}.items()}
    plain_parts = split_state_dict(sd)
    ema_parts = split_state_dict(sd, use_ema=True)
    backbone = exp.model.model if kind == "dyffusion" else exp.model
    want = {k: v.clone() for k, v in backbone.state_dict().items()}
    ema.copy_to(exp.model)  # what ema_scope does before evaluating
    swapped = 0
    for k, v in backbone.state_dict().items():
        assert torch.equal(ema_parts["model"][k], v), k
        assert torch.equal(plain_parts["model"][k], want[k]), k
        swapped += int(not torch.equal(v, want[k]))
    assert swapped == sum(1 for p in backbone.parameters() if p.requires_grad)  # every trainable tensor, no buffer
    if kind == "dyffusion":  # the frozen interpolator has no shadows and is returned as stored
        assert all(torch.equal(ema_parts["interpolator"][k], v) for k, v in interp.model.state_dict().items())
    with pytest.raises(ValueError):
        split_state_dict(exp.state_dict(), use_ema=True)  # a run without use_ema
