"""Writes tests/golden/train_step_kat.json: one training-mode evaluation of the reference's DYffusion objective + backward
(src/diffusion/dyffusion.py:496-567 through torch.autograd; forecaster in train mode = BatchNorm batch statistics, dropout
replaced by the deterministic site masks of oracle/synth.py; frozen interpolator with dropout forced on) on the synthetic
weights / inputs of tests/test_train_oracle_cpu.py.  Stored per case: the loss terms, the L2 norm and a random projection of
every forecaster gradient, and the sums of the updated BatchNorm running statistics.
Build container only:  python tests/golden/make_train_kat.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import test_train_oracle_cpu as T  # noqa: E402

out = {}
for name in T.CASES:
    losses, grads, stats = T.reference_train_step(name)
    out[name] = {**{k: float(v.detach()) for k, v in losses.items()},
                 "grad_norm": {k: float(g.norm()) for k, g in grads.items()},
                 "grad_proj": {k: T.projection(k, g) for k, g in grads.items()},
                 "running": {k: float(v.double().sum()) for k, v in stats.items()}}
    print(name, {k: out[name][k] for k in ("loss", "loss_forward", "loss_forward2")}, len(grads), "gradients")
with open(T.KAT, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
