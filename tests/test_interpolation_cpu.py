"""CPU pinning of `dyffusion_b200.interpolation.InterpolationEvaluation` (host logic around an interpolator backbone) against
the reference's `InterpolationExperiment._evaluation_step` / `get_loss` (src/experiment_types/interpolation.py:68-167), run in
the build container through the shims: the mirror drives the REFERENCE's own backbone module, so with equal seeds (dropout
on -- interpolation runs evaluate with inference dropout) every tensor must be equal bit for bit."""
import pytest
import torch

from oracle import configs as C
from tests import helpers as H

pytestmark = pytest.mark.needs_reference


def _build(dataset, horizon, members):
    from oracle import ref_build
    from tests.golden.make_golden import load_synth
    exp = ref_build.build_interpolator(dataset, horizon=horizon)
    load_synth(exp.model, seed=2)
    exp.hparams.num_predictions = members
    exp.hparams.enable_inference_dropout = True
    return exp


@pytest.mark.parametrize("dataset,horizon,members,batch", [("spring", 4, 3, 2), ("spring", 3, 1, 3), ("sst", 3, 2, 2)])
def test_evaluation_step_equals_reference(dataset, horizon, members, batch):
    from dyffusion_b200.interpolation import InterpolationEvaluation
    exp = _build(dataset, horizon, members)
    d = C.DATASETS[dataset]
    data = {"dynamics": H.synth_tensor(f"ipol.{dataset}.dyn", (batch, 1 + horizon, d["channels"], *d["spatial"]))}
    if d["static"]:
        data["condition"] = H.synth_tensor(f"ipol.{dataset}.static", (batch, d["static"], *d["spatial"]), kind="mask")
    torch.manual_seed(7)
    want = exp.evaluation_step(dict(data), 0, "predict", return_only_preds_and_targets=True)  # incl. the dropout scope
    mine = InterpolationEvaluation(exp.model, horizon=horizon, num_predictions=members)
    torch.manual_seed(7)
    got = mine.evaluation_step(data, "predict")
    assert list(got) == list(want) == [f"t{t}_{s}" for t in range(1, horizon) for s in ("preds", "targets")]
    for k in want:
        assert torch.equal(got[k], want[k]), k
    if members > 1:
        assert tuple(got["t1_preds"].shape) == (members, batch, d["channels"], *d["spatial"])
        assert not torch.equal(got["t1_preds"][0], got["t1_preds"][1])  # inference dropout is on: members differ


def test_training_batch_equals_reference():
    from dyffusion_b200.interpolation import InterpolationEvaluation
    exp = _build("spring", 5, 1)
    data = {"dynamics": H.synth_tensor("ipol.loss.dyn", (4, 6, 4, 10, 10)),
            "condition": H.synth_tensor("ipol.loss.static", (4, 1, 10, 10), kind="mask")}
    exp.train()
    torch.manual_seed(9)
    want = exp.get_loss(dict(data))
    mine = InterpolationEvaluation(exp.model, horizon=5)
    torch.manual_seed(9)
    got = mine.get_loss(data)
    assert float(got.detach()) == float(want.detach())
    with pytest.raises(AssertionError):
        InterpolationEvaluation(exp.model, horizon=1)
    with pytest.raises(AssertionError):
        mine.get_inputs_from_dynamics(data["dynamics"][:, :4])
