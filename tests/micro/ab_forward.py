"""Developer tool (GPU): run one NS forward (3 rows, seeded) and save / compare the output -- for A/B checks of kernel
variants selected by process-wide environment switches (e.g. DYF_UMMA_A=cpasync vs the TMA patch path)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from oracle.synth import synth_tensor  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.gpu_helpers import build_backbone  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--save")
ap.add_argument("--compare")
ap.add_argument("--dataset", default="ns")
ap.add_argument("--rows", type=int, default=3)
a = ap.parse_args()
net = build_backbone(a.dataset, "I", seed=1)
rows = a.rows
x, c = H.forward_inputs(a.dataset, "I", rows=rows)
t = torch.linspace(1.0, 3.0, rows).cuda()
with torch.no_grad():
    y = net(x.cuda(), time=t, condition=None if c is None else c.cuda()).cpu()
if a.save:
    torch.save(y, a.save)
if a.compare:
    y0 = torch.load(a.compare)
    print(f"A/B {a.dataset}: rel-L2 {H.rel_l2(y, y0):.3e}  bit-equal {torch.equal(y, y0)}  finite {bool(torch.isfinite(y).all())}")
