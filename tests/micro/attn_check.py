"""Developer tool (GPU): the fused linear-attention kernels (attn_fused.cu) against the un-fused path
(DYF_DISABLE_ATTN_FUSE=1 at net creation) and the oracle on SST forwards; then per-class timing of a 304-row forward."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

import dyffusion_b200.engine as E  # noqa: E402
from oracle.synth import synth_state_dict  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_backbone  # noqa: E402

SHAPES = H.golden_json("state_shapes.json")
for role in ("F", "I"):
    os.environ["DYF_DISABLE_ATTN_FUSE"] = "1"
    old = build_backbone("sst", role, seed=1)
    del os.environ["DYF_DISABLE_ATTN_FUSE"]
    new = build_backbone("sst", role, seed=1)
    for rows in (1, 3):
        x, cond = H.forward_inputs("sst", role, rows=rows)
        t = torch.linspace(0.5, 3.0, rows)
        c = None if cond is None else cond.cuda()
        with torch.no_grad():
            yo = old(x.cuda(), time=t.cuda(), condition=c)
            yn = new(x.cuda(), time=t.cuda(), condition=c)
            torch.cuda.synchronize()
            sd = synth_state_dict(SHAPES[f"sst_{role}"], seed=1)
            yr = H.oracle_net("sst", role, sd)(x, t, cond)
            torch.manual_seed(5); old._drop_stream = 0
            with old.inference_dropout_scope(True):
                do = old(x.cuda(), time=t.cuda(), condition=c)
            torch.manual_seed(5); new._drop_stream = 0
            with new.inference_dropout_scope(True):
                dn = new(x.cuda(), time=t.cuda(), condition=c)
        print(f"{role} rows={rows}: fused vs unfused {H.rel_l2(yn.cpu(), yo.cpu()):.2e}  fused vs oracle {H.rel_l2(yn.cpu(), yr):.2e}  "
              f"unfused vs oracle {H.rel_l2(yo.cpu(), yr):.2e}  with dropout (same masks) {H.rel_l2(dn.cpu(), do.cpu()):.2e}", flush=True)

rows = 304
for name, net in (("unfused", old), ("fused", new)):
    x, cond = H.forward_inputs("sst", "I", rows=rows)
    x, t = x.cuda(), torch.full((rows,), 2.0).cuda()
    with torch.no_grad(), net.inference_dropout_scope(True):
        for _ in range(2):
            net(x, time=t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            net(x, time=t)
        e1.record()
        torch.cuda.synchronize()
        E.profile_enable(True)
        net(x, time=t)
        torch.cuda.synchronize()
        prof = E.profile_read()
        E.profile_enable(False)
    print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms per 304-row forward ",
          {k: (round(v['ms'], 2), v['launches']) for k, v in prof.items() if v['launches']}, flush=True)
