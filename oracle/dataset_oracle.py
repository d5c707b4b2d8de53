"""TEST INFRASTRUCTURE ONLY -- numpy restatement of `PhysicalSystemsBenchmarkDataModule.create_dataset_multi_horizon`
(`src/datamodules/physical_systems_benchmark.py:191-243`), the checker for `dyffusion_b200.datasets` / `dyf_window_gather`
(second half).  Only tests/ may import this.

Pinned: tests/test_dataset_cpu.py calls the reference method itself (through oracle/ref_shims.py, build container) on
synthetic trajectory objects of the shape `TrajectoryDataset.__getitem__` returns
(src/datamodules/datasets/physical_systems_benchmark.py:31-160) and compares bit for bit."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np


def create_dataset_multi_horizon(trajectories: Sequence, window: int, horizon: int,
                                 num_trajectories: Optional[int] = None) -> Dict[str, object]:
    """trajectories[i] has `.features` (T_i, C, H, W), `.condition` (Cs, H, W) and `.trajectory_meta["num_time_steps"]`.
    -> {"dynamics": (examples, window + horizon, C, H, W), "condition": (examples, Cs, H, W), "origin": [(i, offset)]}."""
    n = len(trajectories) if num_trajectories is None else min(len(trajectories), num_trajectories)  # :205-207
    dyn, cond, origin = [], [], []
    for i in range(n):
        tr = trajectories[i]
        traj_len = tr.trajectory_meta["num_time_steps"]
        time_len = traj_len - horizon - window + 1                                     # :211
        feats = tr.features
        assert feats.shape[0] == traj_len
        view = np.lib.stride_tricks.sliding_window_view(feats, time_len, axis=0)       # :219 (window+horizon, C, H, W, example)
        dyn.append(np.moveaxis(view, -1, 0))                                           # :220 "horizon c h w example -> example horizon c h w"
        cond.append(np.repeat(np.expand_dims(tr.condition, axis=0), time_len, axis=0))  # :216
        origin += [(i, e) for e in range(time_len)]
    return {"dynamics": np.concatenate(dyn, axis=0), "condition": np.concatenate(cond, axis=0), "origin": origin}
