"""Profiling driver: two Navier-Stokes interpolator forwards (dropout on) through the engine; used under ncu."""
import sys
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from tests.gpu_helpers import build_backbone
from oracle.synth import synth_tensor
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8
net = build_backbone("ns", "I", seed=1)
x = synth_tensor("p.x", (rows, 6, 221, 42)).cuda(); c = synth_tensor("p.c", (rows, 2, 221, 42), kind="mask").cuda()
t = torch.full((rows,), 3.0).cuda()
with torch.no_grad(), net.inference_dropout_scope(True):
    for _ in range(2):
        y = net(x, time=t, condition=c)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
