"""Profiling driver: SST interpolator forwards (dropout on) through the engine; used under ncu."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from oracle.synth import synth_tensor  # noqa: E402
from tests.gpu_helpers import build_backbone  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 304
net = build_backbone("sst", "I", seed=1)
x = synth_tensor("p.x", (rows, 2, 60, 60)).cuda()
t = torch.full((rows,), 3.0).cuda()
with torch.no_grad(), net.inference_dropout_scope(True):
    for _ in range(2):
        y = net(x, time=t)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
