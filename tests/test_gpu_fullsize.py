"""Full-size checks of the BASELINE.json configurations through size-independent properties (the oracle would need hours
at these sizes): the property the multi-GPU sharding rests on -- a row's forecast does not depend on which other rows share
its batch -- checked bit for bit between the full batch and a small slice of it, plus finiteness, output keys
and run-to-run reproducibility.  Interpolator dropout is switched off for the comparison (its Philox stream is keyed by the
row index inside the call, so a slice would draw other masks by design).

configs[1] Navier-Stokes 221x42, h=16, 64 rows;  configs[2] SST 60x60, h=7 (+k=25 -> 32 steps), 304 rows = 8 ranks x 38
(forward_conditioning "data" here: "data+noise" draws per-call noise);  configs[3] spring-mesh h=134, refinement, 200 rows."""
import pytest
import torch

from tests import helpers as H
from tests.gpu_helpers import build_dyffusion

pytestmark = pytest.mark.gpu

CASES = [("ns", 64, {}, 16), ("sst", 304, dict(forward_conditioning="data"), 7), ("spring", 200, {}, 134)]


@pytest.mark.parametrize("dataset,rows,overrides,n_keys", CASES)
def test_full_size_sampling_rows_are_independent(dataset, rows, overrides, n_keys):
    dyf = build_dyffusion(dataset, enable_interpolator_dropout=False, **overrides)
    ic, static = H.sampler_case_inputs(f"full.{dataset}", dataset, rows)
    ic = ic.cuda()
    static = None if static is None else static.cuda()
    lo, hi = rows - 5, rows - 2  # a slice from the tail of the batch: the last work items of every persistent kernel
    with torch.no_grad():
        full = dyf.sample(ic, static_condition=static)
        part = dyf.sample(ic[lo:hi].contiguous(), static_condition=None if static is None else static[lo:hi].contiguous())
        again = dyf.sample(ic, static_condition=static)
    torch.cuda.synchronize()
    assert sorted(full) == sorted(f"t{i}_preds" for i in range(1, n_keys + 1))
    for k, v in full.items():
        assert tuple(v.shape) == (rows, *ic.shape[1:]) and torch.isfinite(v).all(), k
        assert torch.equal(v[lo:hi], part[k]), k      # what a rank computes on its shard is what the full batch computes
        assert torch.equal(v, again[k]), k            # bit-reproducible
