"""Developer tool (GPU): times the Navier-Stokes test-time rollout of the reference's shipped setting (horizon 16,
prediction horizon 64 = 4 autoregressive sampler calls, boundary conditions after every horizon) through
`dyffusion_b200.rollout.MultiHorizonRollout`, next to 4 bare `sample()` calls on the same rows, and the same rollout with
the per-horizon device->host copies the reference does (`torch_to_numpy`, forecasting_multi_horizon.py:185-187)."""
import argparse
import functools
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from oracle.synth import synth_tensor  # noqa: E402
from tests.gpu_helpers import build_dyffusion  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=16)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--ar-steps", type=int, default=3)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
from dyffusion_b200.boundary import boundary_conditions  # noqa: E402
from dyffusion_b200.rollout import MultiHorizonRollout  # noqa: E402

h, b, n = 16, a.batch, a.members
dyf = build_dyffusion("ns", horizon=h)
T = 1 + h * (a.ar_steps + 1)
batch = {"dynamics": synth_tensor("rt.dyn", (b, T, 3, 221, 42)).cuda(),
         "condition": synth_tensor("rt.static", (b, 2, 221, 42), kind="mask").cuda(),
         "metadata": {"fixed_mask": (synth_tensor("rt.fixed", (b, 3, 221, 42), kind="mask") > 0).cuda(),
                      "vertices": (synth_tensor("rt.vert", (b, 2, 221, 42)).abs() * 0.2).cuda(),
                      "in_velocity": (0.5 + synth_tensor("rt.vel", (b, 1)).abs()).cuda()}}
t0, dt = torch.zeros(b).cuda(), torch.full((b,), 0.05).cuda()
ro = MultiHorizonRollout(dyf, horizon=h, num_predictions=n, autoregressive_steps=a.ar_steps)
bc = functools.partial(boundary_conditions, "navier-stokes")
ic = ro.transform_inputs(batch["dynamics"][:, :1], split="test").contiguous()
static = ro.get_ensemble_inputs(batch["condition"], "test", add_noise=False).contiguous()


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


with torch.no_grad():
    ms_bare = timed(lambda: [dyf.sample(ic, static_condition=static) for _ in range(a.ar_steps + 1)])
    ms_roll = timed(lambda: ro.evaluation_step(batch, "test", boundary_conditions=bc, t0=t0, dt=dt))
    ms_numpy = timed(lambda: ro.evaluation_step(batch, "test", boundary_conditions=bc, t0=t0, dt=dt, to_numpy=True))
    ms_test = timed(lambda: ro.test_step(batch, boundary_conditions=bc, t0=t0, dt=dt))
rows = n * b
cells = rows * 221 * 42 * h * (a.ar_steps + 1)
print(json.dumps({"rows": rows, "members": n, "batch": b, "sampler_calls": a.ar_steps + 1, "horizons": h * (a.ar_steps + 1),
                  "ms_bare_sample_calls": round(ms_bare, 2), "ms_rollout_device": round(ms_roll, 2),
                  "ms_rollout_numpy_per_horizon": round(ms_numpy, 2), "ms_test_step_with_metrics": round(ms_test, 2),
                  "rollout_Mcellsteps_per_s": round(cells / ms_roll / 1e3, 2)}))
