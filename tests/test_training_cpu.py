"""CPU tests of the DYffusion objective (forward half): `p_losses` of the drop-in class
(dyffusion_b200/diffusion/dyffusion.py) and of the oracle (oracle/dyffusion_oracle.py::p_losses) against the reference's
own `DYffusion.p_losses` (src/diffusion/dyffusion.py:496-567), run in the build container through the shims.

The drop-in's host logic only calls `model.predict_forward` / `interpolator.predict`, so here it is wrapped around the
REFERENCE's torch backbones: losses AND gradients must equal the reference's bit for bit, dropout on (same RNG stream)."""
import json
import os

import pytest
import torch

from oracle import configs as C
from oracle import dyffusion_oracle as O
from tests import helpers as H

KAT = os.path.join(H.GOLDEN, "p_losses_kat.json")
CASES = {
    # name: dataset, horizon, diffusion overrides, per-row diffusion steps
    "spring_h4": ("spring", 4, {}, [0, 1, 3, 2, 0]),
    "spring_h4_lam2_0": ("spring", 4, dict(lambda_reconstruction=1.0, lambda_reconstruction2=0.0), [3, 3, 1]),
    "spring_h3_all_last": ("spring", 3, {}, [2, 2]),
    "sst_h3_k2": ("sst", 3, dict(additional_interpolation_steps=2), [0, 4, 1, 3]),
}


def _inputs(name, dataset, rows):
    d = C.DATASETS[dataset]
    Hh, Ww = d["spatial"]
    cond = H.synth_tensor(f"{name}.cond", (rows, d["channels"], Hh, Ww))
    last = H.synth_tensor(f"{name}.last", (rows, d["channels"], Hh, Ww))
    static = H.synth_tensor(f"{name}.static", (rows, d["static"], Hh, Ww), kind="mask") if d["static"] else None
    return last, cond, static


def _oracle_losses(name):
    dataset, horizon, ov, steps = CASES[name]
    shapes = H.golden_json("state_shapes.json")
    sdF = H.synth_state_dict({k: tuple(v) for k, v in shapes[f"{dataset}_F"].items()}, seed=3)
    sdI = H.synth_state_dict({k: tuple(v) for k, v in shapes[f"{dataset}_I"].items()}, seed=2)
    dk = C.diffusion_kwargs(dataset, horizon=horizon, **ov)
    last, cond, static = _inputs(name, dataset, len(steps))
    counter = {"n": 0}

    def noise(x):
        counter["n"] += 1
        return H.synth_tensor(f"{name}.noise{counter['n'] - 1}", tuple(x.shape))

    with torch.no_grad():
        return O.p_losses(H.oracle_net(dataset, "F", sdF), H.oracle_net(dataset, "I", sdI), H.oracle_schedule(dk), last, cond,
                          torch.tensor(steps), static, forward_conditioning=dk["forward_conditioning"],
                          lambda_reconstruction=dk["lambda_reconstruction"],
                          lambda_reconstruction2=dk["lambda_reconstruction2"], noise_fn=noise)


def _reference(name, dropout):
    from oracle import ref_build
    from tests.golden.make_golden import load_synth
    dataset, horizon, ov, steps = CASES[name]
    ipol = ref_build.build_interpolator(dataset, horizon=horizon)
    exp = ref_build.build_dyffusion(dataset, ipol, horizon=horizon, enable_interpolator_dropout=dropout,
                                    loss_function="l1", **ov)
    load_synth(ipol.model, seed=2), load_synth(exp.model.model, seed=3)
    return exp, ipol, C.diffusion_kwargs(dataset, horizon=horizon, enable_interpolator_dropout=dropout, **ov)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_p_losses_known_answers(name):
    """Committed values = the reference's own p_losses (validation mode, interpolator dropout off) on the synthetic weights."""
    with open(KAT) as f:
        want = json.load(f)[name]
    got = _oracle_losses(name)
    for k in ("loss", "loss_forward", "loss_forward2"):
        assert abs(float(got[k]) - want[k]) <= 2e-5 * max(1.0, abs(want[k])), (name, k, float(got[k]), want[k])


@pytest.mark.needs_reference
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_p_losses_equals_reference(name):
    dataset, horizon, ov, steps = CASES[name]
    exp, ipol, dk = _reference(name, dropout=False)
    last, cond, static = _inputs(name, dataset, len(steps))
    real, counter = torch.randn_like, {"n": 0}

    def fake(x):
        counter["n"] += 1
        return H.synth_tensor(f"{name}.noise{counter['n'] - 1}", tuple(x.shape))

    torch.randn_like = fake
    try:
        with torch.no_grad():
            want = exp.model.p_losses(last, cond, torch.tensor(steps), static_condition=static)
    finally:
        torch.randn_like = real
    got = _oracle_losses(name)
    for k, rk in (("loss", "loss"), ("loss_forward", "val/loss_forward"), ("loss_forward2", "val/loss_forward2")):
        assert abs(float(got[k]) - float(want[rk])) <= 2e-5 * max(1.0, abs(float(want[rk]))), (k, float(got[k]), float(want[rk]))


@pytest.mark.needs_reference
@pytest.mark.parametrize("name,train", [("spring_h4", True), ("spring_h4", False), ("spring_h4_lam2_0", True),
                                        ("spring_h3_all_last", True), ("sst_h3_k2", True)])
def test_dropin_p_losses_equals_reference_bit_for_bit(name, train):
    from dyffusion_b200.diffusion import DYffusion
    dataset, horizon, ov, steps = CASES[name]
    exp, ipol, dk = _reference(name, dropout=True)
    ref = exp.model
    mine = DYffusion(model=ref.model, interpolator=ipol, loss_function="l1", verbose=False, **dk)
    last, cond, static = _inputs(name, dataset, len(steps))
    t = torch.tensor(steps)
    outs = []
    for diff in (ref, mine):
        diff.train(train)
        ipol.eval()  # the frozen interpolator stays in eval mode; its dropout is switched by q_sample's scope
        ref.model.zero_grad()
        torch.manual_seed(21)
        with torch.set_grad_enabled(train):
            d = diff.p_losses(last, cond, t, static_condition=static)
            if train:
                d["loss"].backward()
        grads = [p.grad.clone() for p in ref.model.parameters()] if train else []
        outs.append((d, grads))
    (dr, gr), (dm, gm) = outs
    assert list(dr) == list(dm) and ("train/loss_forward" if train else "val/loss_forward") in dm
    val = lambda v: float(v.detach()) if torch.is_tensor(v) else float(v)
    for k in dr:
        assert val(dr[k]) == val(dm[k]), k
    assert len(gr) == len(gm) and all(torch.equal(a, b) for a, b in zip(gr, gm))
    if train:
        assert all(p.grad is None for p in ipol.parameters())  # frozen interpolator (:461-478)
    # `forward` draws the per-row step like the reference (_base_diffusion.py:81-106)
    torch.manual_seed(4)
    a = ref(cond, last, condition=static)
    torch.manual_seed(4)
    b = mine(cond, last, condition=static)
    assert val(a["loss"]) == val(b["loss"])


def test_engine_backbones_refuse_to_train():
    """No backward kernels yet: training through the engine must fail loudly, not fall back."""
    from tests.gpu_helpers import build_dyffusion
    dyf = build_dyffusion("spring", device="cpu", horizon=4)
    dyf.train()
    last, cond, static = _inputs("spring_h4", "spring", 2)
    with pytest.raises((NotImplementedError, RuntimeError)):
        dyf.p_losses(last, cond, torch.tensor([0, 1]), static_condition=static)
