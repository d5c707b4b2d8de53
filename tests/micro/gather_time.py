"""Developer tool (GPU): achieved HBM bandwidth of `dyf_window_gather` on Navier-Stokes examples (64 examples x 17 frames x
27 846 floats = 121 MB gathered per batch; algorithmic bytes = read + write) and on spring-mesh examples."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

import dyffusion_b200.engine as E  # noqa: E402
from dyffusion_b200.datasets import window_gather  # noqa: E402

res = {}
for name, frame, n_frames, L, batch in (("navier-stokes", (3, 221, 42), 20 * 65, 17, 64), ("spring-mesh", (4, 10, 10), 100 * 805, 135, 256),
                                       ("navier-stokes-16B", (4, 221, 42), 20 * 65, 17, 64)):
    store = torch.randn(n_frames, *frame, device="cuda")
    first = torch.randint(0, n_frames - L, (batch,), generator=torch.Generator().manual_seed(0)).tolist()
    for _ in range(3):
        out = window_gather(store, first, L)
    torch.cuda.synchronize()
    E.profile_enable(True)  # CUDA events around the kernel itself: the host side of a call (index table, torch.empty) costs
    for _ in range(20):     # about as much as the 40 us kernel, so back-to-back wall time would measure Python
        out = window_gather(store, first, L)
    torch.cuda.synchronize()
    prof = E.profile_read()["pack"]
    E.profile_enable(False)
    ms = prof["ms"] / 20
    nbytes = 2 * out.numel() * 4
    ref = torch.stack([store[f:f + L] for f in first])
    res[name] = {"ms": round(ms, 4), "MB_moved": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1), "exact": bool(torch.equal(ref, out)), "launches_per_call": prof["launches"] // 20}
print(json.dumps(res))
