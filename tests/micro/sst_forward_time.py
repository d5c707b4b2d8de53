"""Developer tool (GPU): per-class time of one 304-row SST interpolator forward (dropout on) with the loaded library."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

import dyffusion_b200.engine as E  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_backbone  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 304
net = build_backbone("sst", "I", seed=1)
x, _ = H.forward_inputs("sst", "I", rows=rows)
x, t = x.cuda(), torch.full((rows,), 2.0).cuda()
with torch.no_grad(), net.inference_dropout_scope(True):
    for _ in range(3):
        net(x, time=t)
    torch.cuda.synchronize()
    E.profile_enable(True)
    for _ in range(3):
        net(x, time=t)
    torch.cuda.synchronize()
    prof = E.profile_read()
    E.profile_enable(False)
print(os.path.basename(os.environ.get("DYF_LIB", "default")), {k: round(v["ms"] / 3, 3) for k, v in prof.items() if v["launches"]}, flush=True)
