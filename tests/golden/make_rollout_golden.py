"""Generates tests/golden/rollout_*.pt by running the UNMODIFIED reference evaluation loop
(`MultiHorizonForecastingDYffusion._evaluation_step`, src/experiment_types/forecasting_multi_horizon.py:115-238, with the
datamodule's own `boundary_conditions`, src/datamodules/physical_systems_benchmark.py:245-297) through oracle/ref_shims.py on
the synthetic weights / batches of oracle/synth.py + tests/helpers.py.  Build container only:

    python tests/golden/make_rollout_golden.py
"""
from __future__ import annotations

import functools
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_build, ref_shims  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.golden.make_golden import load_synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def reference_rollout(name, dropout=False, seed=None):
    """-> (return_dict of the reference's `_evaluation_step`, the experiment, the batch, t0, dt)."""
    ref_shims.install()
    from src.datamodules.physical_systems_benchmark import PhysicalSystemsBenchmarkDataModule as DM
    c, batch, t0, dt = H.rollout_case(name)
    ipol = ref_build.build_interpolator(c["dataset"], horizon=c["horizon"])
    exp = ref_build.build_dyffusion(c["dataset"], ipol, horizon=c["horizon"], enable_interpolator_dropout=dropout)
    load_synth(ipol.model, seed=2)
    load_synth(exp.model.model, seed=3)
    exp.hparams.num_predictions = c["members"]
    exp.hparams.autoregressive_steps = c["ar_steps"]
    ipol.hparams.num_predictions = c["members"]  # forecasting_multi_horizon.py:396-398
    fake_dm = types.SimpleNamespace(hparams=types.SimpleNamespace(physical_system=c["system"]))
    bc = functools.partial(DM.boundary_conditions, fake_dm)
    if seed is not None:
        torch.manual_seed(seed)
    work = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}  # the reference scales dynamics by 1e6
    out = exp._evaluation_step(work, 0, "test", boundary_conditions=bc, t0=t0, dt=dt)
    return out, exp, bc, batch, t0, dt


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for name in H.ROLLOUT_CASES:
        out, *_ = reference_rollout(name)
        preds = {k: torch.from_numpy(v).clone() for k, v in out.items() if k.endswith("_preds")}
        torch.save({"preds": preds, "case": H.ROLLOUT_CASES[name]}, os.path.join(OUT, f"{name}.pt"))
        print(name, {k: (tuple(v.shape), round(float(v.abs().mean()), 4)) for k, v in preds.items()})


if __name__ == "__main__":
    main()
