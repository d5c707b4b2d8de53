"""GPU parity of the resident example store (`dyffusion_b200.datasets` -> ctypes -> `dyf_window_gather`) against the numpy
oracle of `create_dataset_multi_horizon` (physical_systems_benchmark.py:191-243).  Data movement only: bit-exact."""
import functools

import numpy as np
import pytest
import torch

from oracle import dataset_oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("system,lengths,window,horizon", [
    ("spring-mesh", [12, 9, 15], 1, 4),     # 400-float frames: 16-byte path
    ("spring-mesh", [7, 7], 2, 3),          # window 2
    ("navier-stokes", [6, 5, 9], 1, 2),     # 27 846-float frames: 8-byte path, ragged trajectory lengths
])
def test_batches_equal_the_reference_examples(system, lengths, window, horizon):
    from dyffusion_b200.datasets import TrajectoryWindows
    import dyffusion_b200.engine as E
    trajs = H.synth_trajectories(system, lengths)
    want = dataset_oracle.create_dataset_multi_horizon(trajs, window, horizon)
    ds = TrajectoryWindows(trajs, window, horizon, physical_system=system)
    n = len(ds)
    before = E.launch_count()
    order = torch.randperm(n, generator=torch.Generator().manual_seed(3)).tolist()
    b = ds.get_batch(order)
    assert E.launch_count() > before
    assert np.array_equal(b["dynamics"].cpu().numpy(), want["dynamics"][order])
    assert np.array_equal(b["condition"].cpu().numpy(), want["condition"][order])
    for j, g in enumerate(order):
        i, _ = want["origin"][g]
        assert np.array_equal(b["metadata"]["fixed_mask"][j].cpu().numpy(), trajs[i].fixed_mask)
        assert float(b["metadata"]["time_step_size"][j]) == np.float32(trajs[i].trajectory_meta["time_step_size"])
        if system == "navier-stokes":
            assert np.array_equal(b["metadata"]["vertices"][j].cpu().numpy(), trajs[i].vertices)
            assert float(b["metadata"]["in_velocity"][j]) == np.float32(trajs[i].trajectory_meta["in_velocity"])
        else:
            assert np.array_equal(b["metadata"]["features"][j, 0].cpu().numpy(), trajs[i].features[0])
    # sequential loader: every example once, in the reference's order; the tail batch is short
    got = torch.cat([x["dynamics"] for x in ds.batches(4)]).cpu().numpy()
    assert np.array_equal(got, want["dynamics"])
    assert sum(x["dynamics"].shape[0] for x in ds.batches(4, drop_last=True)) == (n // 4) * 4
    seen = torch.cat([x["dynamics"] for x in ds.batches(5, shuffle=True, generator=torch.Generator().manual_seed(1))])
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).tolist()
    assert np.array_equal(seen.cpu().numpy(), want["dynamics"][perm])


def test_many_examples_per_launch_and_errors():
    from dyffusion_b200.datasets import TrajectoryWindows, window_gather
    import dyffusion_b200.engine as E
    trajs = H.synth_trajectories("spring-mesh", [300, 250])
    want = dataset_oracle.create_dataset_multi_horizon(trajs, 1, 5)
    ds = TrajectoryWindows(trajs, 1, 5, "spring-mesh")
    idx = list(range(len(ds)))[::-1]  # 540 examples: two launches (448 table entries per launch)
    assert np.array_equal(ds.get_batch(idx)["dynamics"].cpu().numpy(), want["dynamics"][idx])
    odd = torch.arange(7 * 3 * 5, dtype=torch.float32).reshape(7, 3, 5).cuda()  # 15-float frames: 4-byte path
    assert torch.equal(window_gather(odd, [4, 0, 2], 3), torch.stack([odd[4:7], odd[0:3], odd[2:5]]))
    with pytest.raises(ValueError):
        window_gather(odd, [5], 3)  # runs past the end of the store
    with pytest.raises(ValueError):
        window_gather(odd, [-1], 3)
    with pytest.raises(E.EngineError):
        window_gather(odd.cpu(), [0], 3)
    assert window_gather(odd, [], 3).shape == (0, 3, 3, 5)


def test_store_to_rollout_to_metrics_on_device():
    """The evaluation pipeline end to end without leaving the device: resident store -> batch -> autoregressive rollout with
    boundary conditions -> stacked trajectory -> ensemble metrics (the reference's test_step, forecasting_multi_horizon.py:240-262)."""
    from dyffusion_b200.boundary import boundary_conditions
    from dyffusion_b200.datasets import TrajectoryWindows
    from dyffusion_b200.rollout import MultiHorizonRollout
    from tests.gpu_helpers import build_dyffusion
    h, ar = 3, 1
    trajs = H.synth_trajectories("spring-mesh", [12, 10])
    ds = TrajectoryWindows(trajs, 1, h * (ar + 1), "spring-mesh")  # test split: horizon = prediction_horizon (get_horizon, :117-121)
    dyf = build_dyffusion("spring", horizon=h, enable_interpolator_dropout=False)
    ro = MultiHorizonRollout(dyf, horizon=h, num_predictions=4, autoregressive_steps=ar)
    bc = functools.partial(boundary_conditions, "spring-mesh")
    batch = ds.get_batch([0, 3, len(ds) - 1])
    kw = ds.boundary_condition_kwargs(batch)
    assert torch.equal(kw["t0"].cpu(), torch.tensor([trajs[0].t[0], trajs[0].t[0], trajs[1].t[0]]))
    out = ro.evaluation_step(batch, "test", boundary_conditions=bc, **kw)
    fm = batch["metadata"]["fixed_mask"]
    base_q = batch["metadata"]["features"][:, 0, 2:]
    for t in range(1, h * (ar + 1) + 1):
        p = out[f"t{t}_preds"]
        assert p.is_cuda and tuple(p.shape) == (4, 3, 4, 10, 10) and torch.isfinite(p).all()
        assert torch.equal(out[f"t{t}_targets"], batch["dynamics"][:, t])
        assert (p[:, :, :2][:, fm[:, :2]] == 0).all()
        assert torch.equal(p[0, :, 2:][fm[:, 2:]], base_q[fm[:, 2:]])
    m = ro.test_step(batch, boundary_conditions=bc, **kw)
    assert all(len(m[k]) == h * (ar + 1) and np.isfinite(m[k]).all() for k in ("crps", "mse", "ssr"))
