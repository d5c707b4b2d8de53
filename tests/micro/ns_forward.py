"""Developer tool (GPU): times one Navier-Stokes backbone forward at the bench batch (64 rows) and, with --check,
compares the tcgen05 path against the mma.sync path and the oracle.  Run under `ncu --metrics gpu__time_duration.sum`
for a per-kernel list.

    python tests/micro/ns_forward.py [--rows 64] [--iters 10] [--role I] [--check]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from oracle.synth import synth_tensor  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_backbone  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--role", default="I")
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import dyffusion_b200.engine as E

    net = build_backbone("ns", a.role, seed=1)
    cin = 6 if a.role == "I" else 3
    x = synth_tensor("mf.x", (a.rows, cin, 221, 42)).cuda()
    c = synth_tensor("mf.c", (a.rows, 2, 221, 42), kind="mask").cuda()
    t = torch.linspace(1.0, 9.0, a.rows).cuda()
    with torch.no_grad():
        for _ in range(3):
            y = net(x, time=t, condition=c)
        torch.cuda.synchronize()
        E.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            y = net(x, time=t, condition=c)
        e1.record()
        torch.cuda.synchronize()
        prof = E.profile_read()
        E.profile_enable(False)
        ms = e0.elapsed_time(e1) / a.iters
        gf = 48.186 if a.role == "I" else 48.161
        print(f"forward rows={a.rows}: {ms:.3f} ms  ({a.rows * gf / ms:.1f} TFLOP/s dense-equivalent)")
        print({k: round(v["ms"] / a.iters, 3) for k, v in prof.items() if v["launches"]})
        if a.check:
            rows = min(a.rows, 2)
            xs, cs, ts = x[:rows], c[:rows], t[:rows]
            y_u = net(xs, time=ts, condition=cs)
            os.environ["DYF_DISABLE_UMMA"] = "1"
            y_m = net(xs, time=ts, condition=cs)
            del os.environ["DYF_DISABLE_UMMA"]
            from oracle.synth import synth_state_dict
            sd = synth_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=1)
            y_o = H.oracle_net("ns", a.role, sd)(xs.cpu(), ts.cpu(), cs.cpu())
            print(f"umma~mma {H.rel_l2(y_u.cpu(), y_m.cpu()):.2e} | umma~oracle {H.rel_l2(y_u.cpu(), y_o):.2e} | "
                  f"mma~oracle {H.rel_l2(y_m.cpu(), y_o):.2e}")


if __name__ == "__main__":
    main()
