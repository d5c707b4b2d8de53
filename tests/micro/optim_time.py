"""Developer tool (GPU): achieved HBM bandwidth of the fused AdamW step (`dyf_adamw_step`: norm pass 4 B/element + update
28 B/element) next to torch.optim.AdamW(foreach) + clip_grad_norm_ on the same ~100-tensor parameter set."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

import dyffusion_b200.engine as E  # noqa: E402
from dyffusion_b200.optim import AdamW  # noqa: E402

torch.manual_seed(0)
shapes = [(512, 512, 3, 3)] * 8 + [(256, 256, 3, 3)] * 16 + [(128, 128, 3, 3)] * 16 + [(64, 64, 3, 3)] * 16 + [(512,)] * 48
hyper = dict(lr=3e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-4)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ps = [torch.nn.Parameter(0.1 * torch.randn(s, device="cuda")) for s in shapes]
n = sum(p.numel() for p in ps)
gs = [torch.randn_like(p) for p in ps]
for p, g in zip(ps, gs):
    p.grad = g
ref = torch.optim.AdamW(ps, foreach=True, **hyper)


def torch_step():
    torch.nn.utils.clip_grad_norm_(ps, 1.0)
    ref.step()


ms_torch = timed(torch_step)
ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
opt = AdamW(ps2, max_grad_norm=1.0, **hyper)
for p, g in zip(ps2, gs):
    p.grad.copy_(g)
ms_wall = timed(opt.step)
E.profile_enable(True)
for _ in range(10):
    opt.step()
torch.cuda.synchronize()
prof = E.profile_read()["elementwise"]
E.profile_enable(False)
ms_kernels = prof["ms"] / 10
print(json.dumps({"elements": n, "tensors": len(ps), "ms_torch_foreach_clip_plus_step": round(ms_torch, 3),
                  "ms_fused_step_wall": round(ms_wall, 3), "ms_fused_kernels": round(ms_kernels, 3),
                  "GBps_fused_kernels": round(32.0 * n / ms_kernels / 1e6, 1), "speedup_vs_torch": round(ms_torch / ms_wall, 2)}))
