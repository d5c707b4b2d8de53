"""Developer tool (GPU): the flat-raster tcgen05 path of the spring-mesh backbone (conv_flat.cu) against the mma.sync path
(DYF_DISABLE_FLAT=1 at net creation) and the oracle, per forward and through the sampler; then timings."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from oracle.synth import synth_state_dict  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_backbone, build_dyffusion  # noqa: E402

SHAPES = H.golden_json("state_shapes.json")


def nets(role):
    os.environ["DYF_DISABLE_FLAT"] = "1"
    old = build_backbone("spring", role, seed=1)
    del os.environ["DYF_DISABLE_FLAT"]
    new = build_backbone("spring", role, seed=1)
    return old, new


for role in ("F", "I"):
    old, new = nets(role)
    for rows in (1, 3, 37):
        x, cond = H.forward_inputs("spring", role, rows=rows)
        t = torch.linspace(0.5, 3.0, rows)
        with torch.no_grad():
            yo = old(x.cuda(), time=t.cuda(), condition=cond.cuda())
            yn = new(x.cuda(), time=t.cuda(), condition=cond.cuda())
            torch.cuda.synchronize()
            sd = synth_state_dict(SHAPES[f"spring_{role}"], seed=1)
            yr = H.oracle_net("spring", role, sd)(x, t, cond)
        print(f"{role} rows={rows}: flat vs mma {H.rel_l2(yn.cpu(), yo.cpu()):.2e}  flat vs oracle {H.rel_l2(yn.cpu(), yr):.2e}  "
              f"mma vs oracle {H.rel_l2(yo.cpu(), yr):.2e}", flush=True)

# sampler: batched logical calls (G = rows, 2 calls per interpolator launch), dropout off and on
for rows in (2, 5):
    os.environ["DYF_DISABLE_FLAT"] = "1"
    a = build_dyffusion("spring", horizon=6, enable_interpolator_dropout=False)
    a_d = build_dyffusion("spring", horizon=6)
    del os.environ["DYF_DISABLE_FLAT"]
    b = build_dyffusion("spring", horizon=6, enable_interpolator_dropout=False)
    b_d = build_dyffusion("spring", horizon=6)
    ic, st = H.sampler_case_inputs("flat", "spring", rows)
    with torch.no_grad():
        oa, ob = a.sample(ic.cuda(), static_condition=st.cuda()), b.sample(ic.cuda(), static_condition=st.cuda())
        errs = [H.rel_l2(ob[k].cpu(), oa[k].cpu()) for k in oa]
        torch.manual_seed(3); a_d._calls = 0
        da = a_d.sample(ic.cuda(), static_condition=st.cuda())
        torch.manual_seed(3); b_d._calls = 0
        db = b_d.sample(ic.cuda(), static_condition=st.cuda())
        derrs = [H.rel_l2(db[k].cpu(), da[k].cpu()) for k in da]
    print(f"sampler rows={rows}: flat vs mma max {max(errs):.2e}; with dropout (same masks) max {max(derrs):.2e}", flush=True)

import dyffusion_b200.engine as E  # noqa: E402
for rows in (200, 800):
    for flat in (False, True):
        if not flat:
            os.environ["DYF_DISABLE_FLAT"] = "1"
        dyf = build_dyffusion("spring")
        os.environ.pop("DYF_DISABLE_FLAT", None)
        ic, st = H.sampler_case_inputs("time", "spring", rows)
        ic, st = ic.cuda(), st.cuda()
        with torch.no_grad():
            for _ in range(3):
                dyf.sample(ic, static_condition=st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                dyf.sample(ic, static_condition=st)
            e1.record()
            torch.cuda.synchronize()
            E.profile_enable(True)
            dyf.sample(ic, static_condition=st)
            torch.cuda.synchronize()
            prof = E.profile_read()
            E.profile_enable(False)
        print(f"rows={rows} flat={flat}: {e0.elapsed_time(e1) / 3:.2f} ms per sample()  ",
              {k: (round(v['ms'], 2), v['launches']) for k, v in prof.items() if v['launches']}, flush=True)
