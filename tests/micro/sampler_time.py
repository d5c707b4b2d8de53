"""Developer tool (GPU): times `sample()` of the shipped SST / spring-mesh / NS DYffusion configs (oracle/configs.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_dyffusion  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dataset", default="sst")
ap.add_argument("--rows", type=int, default=304)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dyf = build_dyffusion(a.dataset)
ic, static = H.sampler_case_inputs("time", a.dataset, a.rows)
ic = ic.cuda()
static = None if static is None else static.cuda()
with torch.no_grad():
    for _ in range(2):
        out = dyf.sample(ic, static_condition=static)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        out = dyf.sample(ic, static_condition=static)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
import dyffusion_b200.engine as E  # noqa: E402
E.profile_enable(True)
with torch.no_grad():
    dyf.sample(ic, static_condition=static)
torch.cuda.synchronize()
prof = E.profile_read()
E.profile_enable(False)
print({k: (round(v["ms"], 2), v["launches"]) for k, v in prof.items() if v["launches"]})
n = len(dyf.sampling_schedule)
hw = ic.shape[-1] * ic.shape[-2]
print(f"{a.dataset}: rows={a.rows} schedule steps={n} outputs={len(out)}  {ms:.2f} ms per sample()  "
      f"{a.rows * hw * n / ms * 1e3 / 1e6:.2f} M cell-steps/s")
