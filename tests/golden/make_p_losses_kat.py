"""Writes tests/golden/p_losses_kat.json: the reference's own `DYffusion.p_losses` (src/diffusion/dyffusion.py:496-567; eval
mode, interpolator dropout off, L1 criterion) on the synthetic weights / inputs of tests/test_training_cpu.py.
Build container only:  python tests/golden/make_p_losses_kat.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import helpers as H  # noqa: E402
from tests import test_training_cpu as T  # noqa: E402

out = {}
for name, (dataset, horizon, ov, steps) in T.CASES.items():
    exp, ipol, dk = T._reference(name, dropout=False)
    last, cond, static = T._inputs(name, dataset, len(steps))
    real, counter = torch.randn_like, {"n": 0}

    def fake(x, _c=counter, _n=name):
        _c["n"] += 1
        return H.synth_tensor(f"{_n}.noise{_c['n'] - 1}", tuple(x.shape))

    torch.randn_like = fake
    try:
        with torch.no_grad():
            d = exp.model.p_losses(last, cond, torch.tensor(steps), static_condition=static)
    finally:
        torch.randn_like = real
    out[name] = {"loss": float(d["loss"]), "loss_forward": float(d["val/loss_forward"]),
                 "loss_forward2": float(d["val/loss_forward2"]), "steps": steps, "noise_draws": counter["n"]}
    print(name, out[name])
with open(T.KAT, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
