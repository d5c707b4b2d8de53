"""Oracle for the training step -- built and pinned BEFORE the backward kernels exist, as the first gate the
next round's CUDA backward has to pass.

`oracle/dyffusion_oracle.py` in train mode (BatchNorm batch statistics + running-statistics update, dropout on, the full
`p_losses` objective) differentiated by torch.autograd, against the REFERENCE modules in train mode differentiated the same
way (build container, through the shims), with the dropout masks of both sides replaced by the deterministic site masks of
oracle/synth.py; and against committed known answers of that reference run (loss terms, per-parameter gradient norms and
random projections, updated running statistics: tests/golden/train_step_kat.json, made by tests/golden/make_train_kat.py)."""
import json
import os

import pytest
import torch

from oracle import configs as C
from oracle import dyffusion_oracle as O
from oracle.synth import SiteDropout, synth_state_dict, synth_tensor
from tests import helpers as H

KAT = os.path.join(H.GOLDEN, "train_step_kat.json")
CASES = {
    # name: dataset, horizon, rows' diffusion steps
    "spring_h4": ("spring", 4, [0, 1, 3, 2, 0, 2]),
    "ns_h3": ("ns", 3, [0, 2, 1]),
    "sst_h3_k2": ("sst", 3, [0, 4, 1, 3]),   # num_timesteps 5 (k = 2 auxiliary steps); "data+noise" forecaster conditioning
}
OVERRIDES = {"sst_h3_k2": dict(additional_interpolation_steps=2)}
DROP_SEED = 5


def inputs(name):
    dataset, horizon, steps = CASES[name]
    d = C.DATASETS[dataset]
    Hh, Ww = d["spatial"]
    r = len(steps)
    cond = synth_tensor(f"{name}.train.cond", (r, d["channels"], Hh, Ww))
    last = synth_tensor(f"{name}.train.last", (r, d["channels"], Hh, Ww))
    static = synth_tensor(f"{name}.train.static", (r, d["static"], Hh, Ww), kind="mask") if d["static"] else None
    return last, cond, static, torch.tensor(steps)


def projection(key, g):
    """A scalar fingerprint of a gradient tensor: its dot product with a fixed synthetic direction."""
    return float((g.double() * synth_tensor(f"proj.{key}", tuple(g.shape)).double()).sum())


def _noise(name):
    """Deterministic stand-in for torch.randn_like ("data+noise" conditioning draws one noise tensor per forecaster call)."""
    counter = {"n": 0}

    def fn(x):
        counter["n"] += 1
        return synth_tensor(f"{name}.train.noise{counter['n'] - 1}", tuple(x.shape))

    return fn


def oracle_train_step(name):
    """-> (loss dict, {param key: grad} of the forecaster, updated BatchNorm running statistics of the forecaster)."""
    dataset, horizon, steps = CASES[name]
    shapes = H.golden_json("state_shapes.json")
    sdF = synth_state_dict({k: tuple(v) for k, v in shapes[f"{dataset}_F"].items()}, seed=3)
    sdI = synth_state_dict({k: tuple(v) for k, v in shapes[f"{dataset}_I"].items()}, seed=2)
    trainable = [k for k, v in sdF.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    for k in trainable:
        sdF[k].requires_grad_(True)
    dk = C.diffusion_kwargs(dataset, horizon=horizon, **OVERRIDES.get(name, {}))
    archF, kwF = H.oracle_kwargs(dataset, "F")
    archI, kwI = H.oracle_kwargs(dataset, "I")
    dropF, dropI, stats = SiteDropout(DROP_SEED), SiteDropout(DROP_SEED + 1), {}
    fF, fI = O.BACKBONES[archF], O.BACKBONES[archI]
    forecaster = lambda x, t, c: fF(sdF, x, t, c, drop=dropF, train_stats=stats, **kwF)   # train mode: batch statistics
    interpolator = lambda x, t, c: fI(sdI, x, t, c, drop=dropI, **kwI)                     # frozen, eval BN, dropout forced on
    last, cond, static, t = inputs(name)
    out = O.p_losses(forecaster, interpolator, H.oracle_schedule(dk), last, cond, t, static,
                     forward_conditioning=dk["forward_conditioning"], lambda_reconstruction=dk["lambda_reconstruction"],
                     lambda_reconstruction2=dk["lambda_reconstruction2"], noise_fn=_noise(name))
    out["loss"].backward()
    return out, {k: sdF[k].grad for k in trainable}, stats


def reference_train_step(name):
    from oracle import ref_build
    from tests.golden.make_golden import load_synth
    dataset, horizon, steps = CASES[name]
    ipol = ref_build.build_interpolator(dataset, horizon=horizon)
    exp = ref_build.build_dyffusion(dataset, ipol, horizon=horizon, loss_function="l1", **OVERRIDES.get(name, {}))
    load_synth(ipol.model, seed=2), load_synth(exp.model.model, seed=3)
    diff = exp.model
    diff.train()
    ipol.eval()
    hF, hI = ref_build.HookedDropout(diff.model, seed=DROP_SEED), ref_build.HookedDropout(ipol.model, seed=DROP_SEED + 1)
    last, cond, static, t = inputs(name)
    real = torch.randn_like
    torch.randn_like = _noise(name)
    try:
        out = diff.p_losses(last, cond, t, static_condition=static)
        out["loss"].backward()
    finally:
        torch.randn_like = real
        hF.remove(), hI.remove()
    grads = {k: p.grad for k, p in diff.model.named_parameters()}
    stats = {k: v.clone() for k, v in diff.model.state_dict().items() if k.endswith(("running_mean", "running_var"))}
    return {"loss": out["loss"], "loss_forward": out["train/loss_forward"], "loss_forward2": out["train/loss_forward2"]}, grads, stats


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_train_step_known_answers(name):
    with open(KAT) as f:
        want = json.load(f)[name]
    out, grads, stats = oracle_train_step(name)
    for k in ("loss", "loss_forward", "loss_forward2"):
        assert abs(float(out[k].detach()) - want[k]) <= 2e-5 * max(1.0, abs(want[k])), (k, float(out[k].detach()), want[k])
    assert sorted(grads) == sorted(want["grad_norm"])
    total = sum(v ** 2 for v in want["grad_norm"].values()) ** 0.5
    for k, g in grads.items():
        assert g is not None, k
        assert abs(float(g.norm()) - want["grad_norm"][k]) <= 1e-3 * want["grad_norm"][k] + 1e-6 * total, k
        assert abs(projection(k, g) - want["grad_proj"][k]) <= 2e-3 * want["grad_norm"][k] * g.numel() ** 0.5 + 1e-6 * total, k
    assert sorted(stats) == sorted(want["running"])
    for k, v in stats.items():
        assert abs(float(v.double().sum()) - want["running"][k]) <= 1e-4 * max(1.0, abs(want["running"][k])), k


@pytest.mark.needs_reference
@pytest.mark.parametrize("name", ["spring_h4", "sst_h3_k2"])
def test_oracle_train_step_equals_reference_autograd(name):
    out, grads, stats = oracle_train_step(name)
    r_out, r_grads, r_stats = reference_train_step(name)
    for k in out:
        assert abs(float(out[k].detach()) - float(r_out[k].detach())) <= 2e-5 * max(1.0, abs(float(r_out[k].detach()))), k
    assert sorted(grads) == sorted(r_grads)
    for k in grads:
        assert H.rel_l2(grads[k], r_grads[k]) <= 1e-4, (k, H.rel_l2(grads[k], r_grads[k]))  # measured: 0.0 (spring-mesh)
    for k in stats:
        assert torch.allclose(stats[k], r_stats[k], rtol=1e-5, atol=1e-6), k
