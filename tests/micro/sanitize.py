"""Developer tool (GPU): one small invocation of every hot-path kernel family, for compute-sanitizer
(`compute-sanitizer --tool memcheck|racecheck python tests/micro/sanitize.py`): a Navier-Stokes forward and a 2-step
sample(), an SST forward with dropout (fused attention / GroupNorm kernels) and a spring-mesh sample() (flat-raster path)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_backbone, build_dyffusion  # noqa: E402

which = sys.argv[1:] or ["ns", "sst", "spring"]
with torch.no_grad():
    if "ns" in which:
        net = build_backbone("ns", "I", seed=1)
        x, c = H.forward_inputs("ns", "I", rows=1)
        with net.inference_dropout_scope(True):
            y = net(x.cuda(), time=torch.tensor([2.0]).cuda(), condition=c.cuda())
        torch.cuda.synchronize()
        print("ns forward ok", float(y.abs().mean()), flush=True)
        dyf = build_dyffusion("ns", horizon=2, cuda_graph=False)
        ic, st = H.sampler_case_inputs("san", "ns", 1)
        out = dyf.sample(ic.cuda(), static_condition=st.cuda())
        torch.cuda.synchronize()
        print("ns sample ok", sorted(out), flush=True)
    if "sst" in which:
        net = build_backbone("sst", "I", seed=1)
        x, _ = H.forward_inputs("sst", "I", rows=2)
        with net.inference_dropout_scope(True):
            y = net(x.cuda(), time=torch.tensor([1.0, 2.5]).cuda())
        torch.cuda.synchronize()
        print("sst forward ok", float(y.abs().mean()), flush=True)
    if "spring" in which:
        dyf = build_dyffusion("spring", horizon=4, cuda_graph=False)
        ic, st = H.sampler_case_inputs("san", "spring", 3)
        out = dyf.sample(ic.cuda(), static_condition=st.cuda())
        torch.cuda.synchronize()
        print("spring sample ok", sorted(out), flush=True)
