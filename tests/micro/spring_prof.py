"""Developer tool (GPU): a few spring-mesh sampler steps for ncu (short horizon, shipped networks)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_dyffusion  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 800
dyf = build_dyffusion("spring", horizon=6, cuda_graph=False)
ic, st = H.sampler_case_inputs("prof", "spring", rows)
with torch.no_grad():
    for _ in range(2):
        dyf.sample(ic.cuda(), static_condition=st.cuda())
torch.cuda.synchronize()
