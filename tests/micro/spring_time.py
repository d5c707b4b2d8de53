"""Developer tool (GPU): spring-mesh sample() time at the shipped config for the given row counts (env knobs apply)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

import dyffusion_b200.engine as E  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_dyffusion  # noqa: E402

tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("DYF_"))
for rows in [int(a) for a in sys.argv[1:]] or [200, 800]:
    dyf = build_dyffusion("spring")
    ic, st = H.sampler_case_inputs("time", "spring", rows)
    ic, st = ic.cuda(), st.cuda()
    with torch.no_grad():
        for _ in range(3):
            dyf.sample(ic, static_condition=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dyf.sample(ic, static_condition=st)
        e1.record()
        torch.cuda.synchronize()
        E.profile_enable(True)
        dyf.sample(ic, static_condition=st)
        torch.cuda.synchronize()
        prof = E.profile_read()
        E.profile_enable(False)
    print(f"[{tag}] rows={rows}: {e0.elapsed_time(e1) / 5:.2f} ms per sample()  ",
          {k: (round(v['ms'], 2), v['launches']) for k, v in prof.items() if v['launches']}, flush=True)
