"""GPU parity of `dyf_ensemble_metrics` (through dyffusion_b200.metrics.evaluate_ensemble_prediction -> ctypes -> C ABI)
against the oracle restatement of src/utilities/evaluation.py.  fp32 inputs, fp32 per-element arithmetic, float64
reductions: tolerance rel 2e-5 on every metric; repeated calls are bit-identical (fixed-order reductions)."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M
from oracle.synth import synth_tensor

pytestmark = pytest.mark.gpu
RTOL = 2e-5


@pytest.mark.parametrize("shape", [(5, 4, 3, 10, 10), (50, 6, 1, 60, 60), (2, 3, 4, 221, 42), (1, 2, 3, 9, 7), (7, 5, 13)])
def test_metrics_match_oracle(shape):
    from dyffusion_b200.metrics import evaluate_ensemble_prediction
    preds = synth_tensor(f"met.p{shape}", shape)
    tg = synth_tensor(f"met.t{shape}", shape[1:]) * 0.7 + 0.1
    if len(shape) == 5:
        preds[:, 0, 0, :2] = preds[0, 0, 0, :2].clone()  # ties inside the ensemble
        tg[1, 0, 0, :3] = preds[min(1, shape[0] - 1), 1, 0, 0, :3].clone()  # observation equal to a member
    for kw in (dict(), dict(mean_over_samples=False), dict(also_per_member_metrics=True)):
        got = evaluate_ensemble_prediction(preds.cuda(), tg.cuda(), **kw)
        ref = M.evaluate_ensemble_prediction(preds.numpy(), tg.numpy(), **kw)
        assert set(got) == set(ref)
        for k in ref:
            np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64),
                                       rtol=RTOL, atol=1e-7, err_msg=f"{k} {kw}")


def test_metrics_are_bit_reproducible_and_need_cuda():
    from dyffusion_b200.metrics import evaluate_ensemble_prediction
    import dyffusion_b200.engine as E
    preds = synth_tensor("met.rep", (8, 3, 3, 221, 42)).cuda()
    tg = synth_tensor("met.rep.t", (3, 3, 221, 42)).cuda()
    a = evaluate_ensemble_prediction(preds, tg, mean_over_samples=False)
    b = evaluate_ensemble_prediction(preds, tg, mean_over_samples=False)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    with pytest.raises(E.EngineError):
        evaluate_ensemble_prediction(preds.cpu(), tg.cpu())
