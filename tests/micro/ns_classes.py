"""Developer tool (GPU): per-kernel-class device time of one Navier-Stokes interpolator forward (CUDA events around every
launch), for A/B runs with the DYF_* switches.  usage: ns_classes.py [rows] [repeats]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from dyffusion_b200 import engine as E  # noqa: E402
from oracle.synth import synth_tensor  # noqa: E402
from tests.gpu_helpers import build_backbone  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
net = build_backbone("ns", "I", seed=1)
x = synth_tensor("p.x", (rows, 6, 221, 42)).cuda()
c = synth_tensor("p.c", (rows, 2, 221, 42), kind="mask").cuda()
t = torch.full((rows,), 3.0).cuda()
with torch.no_grad(), net.inference_dropout_scope(True):
    for _ in range(3):
        net(x, time=t, condition=c)
    torch.cuda.synchronize()
    E.profile_filter(None)
    E.profile_enable(True)
    for _ in range(reps):
        net(x, time=t, condition=c)
    torch.cuda.synchronize()
    prof = E.profile_read()
    E.profile_enable(False)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        net(x, time=t, condition=c)
    t1.record()
    torch.cuda.synchronize()
print(f"rows={rows}: {t0.elapsed_time(t1) / reps * 1e3:.0f} us per forward; per class (us): " +
      "  ".join(f"{k}={v['ms'] / reps * 1e3:.1f}/{v['launches'] // reps}" for k, v in prof.items() if v["launches"]))
