"""CPU tests of the sliding-window example store (second half): the numpy oracle against the reference's
`create_dataset_multi_horizon` itself (build container), and the product's host-side index arithmetic
(`dyffusion_b200.datasets.TrajectoryWindows`: example numbering, start frames) against the oracle.  The gather itself is
CUDA-only (tests/test_gpu_dataset.py)."""
import types

import numpy as np
import pytest
import torch

from oracle import dataset_oracle, ref_shims
from tests import helpers as H

from dyffusion_b200.datasets import TrajectoryWindows

CASES = [("spring-mesh", [12, 9, 15], 1, 4), ("spring-mesh", [7, 7], 2, 3), ("navier-stokes", [6, 5], 1, 2)]


@pytest.mark.needs_reference
@pytest.mark.parametrize("system,lengths,window,horizon", CASES)
def test_oracle_equals_reference_method(system, lengths, window, horizon):
    ref_shims.install()
    from src.datamodules.physical_systems_benchmark import PhysicalSystemsBenchmarkDataModule as DM
    trajs = H.synth_trajectories(system, lengths)
    fake = types.SimpleNamespace(hparams=types.SimpleNamespace(window=window, horizon=horizon, num_trajectories=None,
                                                               physical_system=system),
                                 get_horizon=lambda split: horizon)
    want = DM.create_dataset_multi_horizon(fake, "val", trajs, keep_trajectory_dim=False)
    got = dataset_oracle.create_dataset_multi_horizon(trajs, window, horizon)
    assert np.array_equal(got["dynamics"], want["dynamics"]) and np.array_equal(got["condition"], want["condition"])
    assert len(want["metadata"]) == len(got["origin"]) == got["dynamics"].shape[0]
    for (i, _), m in zip(got["origin"], want["metadata"]):  # the reference repeats trajectory i's meta dict per example
        assert m["name"] == trajs[i].trajectory_meta["name"]
    fake.hparams.num_trajectories = 1  # the training split's cap (:205-207)
    want1 = DM.create_dataset_multi_horizon(fake, "train", trajs, keep_trajectory_dim=False)
    got1 = dataset_oracle.create_dataset_multi_horizon(trajs, window, horizon, num_trajectories=1)
    assert np.array_equal(got1["dynamics"], want1["dynamics"])


@pytest.mark.parametrize("system,lengths,window,horizon", CASES)
def test_host_index_arithmetic_equals_oracle(system, lengths, window, horizon):
    trajs = H.synth_trajectories(system, lengths)
    want = dataset_oracle.create_dataset_multi_horizon(trajs, window, horizon)
    ds = TrajectoryWindows(trajs, window, horizon, physical_system=system, device="cpu")  # index arithmetic only
    assert len(ds) == want["dynamics"].shape[0] == sum(T - window - horizon + 1 for T in lengths)
    assert [ds.origin(g) for g in range(len(ds))] == want["origin"]
    assert ds.origin(-1) == want["origin"][-1]
    first, tr = ds.first_frames(range(len(ds)))
    store = ds.frames.numpy()
    L = window + horizon
    for g in range(len(ds)):
        assert np.array_equal(store[first[g]:first[g] + L], want["dynamics"][g])
        assert np.array_equal(ds.conditions[tr[g]].numpy(), want["condition"][g])
    with pytest.raises(IndexError):
        ds.origin(len(ds))


def test_errors_and_no_cpu_fallback():
    import dyffusion_b200.engine as E
    trajs = H.synth_trajectories("spring-mesh", [8, 8])
    with pytest.raises(AssertionError):
        TrajectoryWindows(trajs, 0, 3, "spring-mesh", device="cpu")
    with pytest.raises(AssertionError):
        TrajectoryWindows(trajs, 1, 0, "spring-mesh", device="cpu")
    with pytest.raises(NotImplementedError):
        TrajectoryWindows(trajs, 1, 3, "pendulum", device="cpu")
    with pytest.raises(ValueError):
        TrajectoryWindows(trajs, 1, 8, "spring-mesh", device="cpu")  # no example fits
    ds = TrajectoryWindows(trajs, 1, 3, "spring-mesh", device="cpu")
    with pytest.raises(E.EngineError):  # the gather has no CPU path
        ds.get_batch([0, 1])


@pytest.mark.needs_reference
def test_random_stores_against_the_reference_method():
    """Random trajectory counts / lengths / windows / horizons (spring-mesh sized frames): oracle == reference method, and the
    product's index arithmetic addresses exactly the reference's examples."""
    import random
    ref_shims.install()
    from src.datamodules.physical_systems_benchmark import PhysicalSystemsBenchmarkDataModule as DM
    rng = random.Random(3)
    for trial in range(12):
        window, horizon = rng.randint(1, 3), rng.randint(1, 6)
        lengths = [window + horizon + rng.randint(0, 7) for _ in range(rng.randint(1, 4))]
        cap = rng.choice([None, None, 1, 2])
        trajs = H.synth_trajectories("spring-mesh", lengths, seed=trial)
        fake = types.SimpleNamespace(hparams=types.SimpleNamespace(window=window, horizon=horizon, num_trajectories=cap,
                                                                   physical_system="spring-mesh"),
                                     get_horizon=lambda split, h=horizon: h)
        want = DM.create_dataset_multi_horizon(fake, "train", trajs, keep_trajectory_dim=False)
        got = dataset_oracle.create_dataset_multi_horizon(trajs, window, horizon, num_trajectories=cap)
        assert np.array_equal(got["dynamics"], want["dynamics"]) and np.array_equal(got["condition"], want["condition"])
        ds = TrajectoryWindows(trajs, window, horizon, "spring-mesh", num_trajectories=cap, device="cpu")
        assert len(ds) == want["dynamics"].shape[0]
        first, tr = ds.first_frames(range(len(ds)))
        store = ds.frames.numpy()
        for g in rng.sample(range(len(ds)), min(len(ds), 6)):
            assert np.array_equal(store[first[g]:first[g] + window + horizon], want["dynamics"][g])
            assert np.array_equal(ds.conditions[tr[g]].numpy(), want["condition"][g])
            assert want["metadata"][g]["name"] == trajs[tr[g]].trajectory_meta["name"]
