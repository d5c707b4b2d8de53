"""GPU-resident drop-in for the physical-systems benchmark's example store (second half).

The reference builds, on the host, one numpy copy of every example -- `create_dataset_multi_horizon`
(`src/datamodules/physical_systems_benchmark.py:191-243`): a `sliding_window_view` over each trajectory re-arranged to
(example, window + horizon, C, H, W) and concatenated, `condition` repeated per example, the per-trajectory metadata dict
repeated per example -- wraps it in `MyTensorDataset`, and lets a DataLoader collate and Lightning copy each batch to the
device.  Here the trajectories are uploaded ONCE, back to back, and a batch is gathered in HBM by example index
(`dyf_window_gather`): the same example numbering (trajectory-major, offset-minor), the same batch dict
(`dynamics`, `condition`, `metadata`), bit-identical values, no per-batch host work beyond the index table.

`metadata` carries what the evaluation path reads (`boundary_conditions`, `get_boundary_condition_kwargs`,
physical_systems_benchmark.py:245-303): `fixed_mask`, `t`, `time_step_size`, `in_velocity` / `vertices` (Navier-Stokes),
`features` (spring-mesh).  One documented difference: the reference collates the WHOLE trajectory into every example's
`metadata["features"]` although only its first frame is read (`metadata["features"][b, 0, 2:]`, :281); here it holds that first
frame only, shape (B, 1, C, H, W).  No CPU / PyTorch fallback for the gather."""
from __future__ import annotations

import bisect
import ctypes
from typing import Any, Dict, Iterator, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine as E


def _get(tr, name, default=None):
    if isinstance(tr, Mapping):
        return tr.get(name, default)
    return getattr(tr, name, default)


def window_gather(store: torch.Tensor, first_frames: Sequence[int], frames_per_example: int) -> torch.Tensor:
    """store: (n_frames, *frame) fp32 CUDA -> (len(first_frames), frames_per_example, *frame)."""
    if not store.is_cuda:
        raise E.EngineError("dyffusion_b200.datasets has no CPU path: the trajectory store must be a CUDA tensor")
    if not (store.is_contiguous() and store.dtype == torch.float32):
        raise ValueError("the trajectory store must be a contiguous float32 tensor")
    n = len(first_frames)
    frame_shape = tuple(store.shape[1:])
    out = torch.empty((n, frames_per_example, *frame_shape), dtype=torch.float32, device=store.device)
    if n == 0:
        return out
    table = (ctypes.c_int64 * n)(*[int(f) for f in first_frames])
    E._check(E.LIB.dyf_window_gather(store.data_ptr(), store.shape[0], int(np.prod(frame_shape, dtype=np.int64)), table, n,
                                     frames_per_example, out.data_ptr(), E._stream_ptr()))
    return out


class TrajectoryWindows:
    """Examples of `window + horizon` consecutive frames over a set of trajectories, numbered as the reference numbers them.

    `trajectories[i]` is a mapping or object with `features` (T_i, C, H, W), `condition` (Cs, H, W) and optionally
    `fixed_mask` (C, H, W), `vertices` (2, H, W), `t` (T_i,), `trajectory_meta` (dict with `num_time_steps`,
    `time_step_size`, `in_velocity`, ...) -- what `TrajectoryDataset.__getitem__` returns
    (src/datamodules/datasets/physical_systems_benchmark.py:66-160)."""

    def __init__(self, trajectories: Sequence[Any], window: int, horizon: int, physical_system: str = "navier-stokes",
                 num_trajectories: Optional[int] = None, device="cuda"):
        assert horizon > 0, f"horizon must be > 0 or a list, but is {horizon}"  # _check_args (:122-126)
        assert window > 0, f"window must be > 0, but is {window}"
        if physical_system not in ("navier-stokes", "spring-mesh"):
            raise NotImplementedError(f"Physical system {physical_system} is not implemented yet.")
        self.window, self.horizon, self.physical_system = int(window), int(horizon), physical_system
        self.frames_per_example = self.window + self.horizon
        n = len(trajectories) if num_trajectories is None else min(len(trajectories), num_trajectories)
        self.device = torch.device(device)
        feats, conds, masks, verts, t_all, self._meta = [], [], [], [], [], []
        self._base: List[int] = []        # first frame of trajectory i in the store
        self._first_example: List[int] = []  # global index of its first example
        self._examples: List[int] = []    # its number of examples (`time_len`, :211)
        frame, ex = 0, 0
        for i in range(n):
            tr = trajectories[i]
            f = np.asarray(_get(tr, "features"), dtype=np.float32)
            meta = dict(_get(tr, "trajectory_meta", {}) or {})
            traj_len = int(meta.get("num_time_steps", f.shape[0]))
            if f.shape[0] != traj_len:  # raise_if_invalid_shape (:214)
                raise ValueError(f"dynamics_i: expected {traj_len} time steps, got {f.shape[0]}")
            time_len = traj_len - self.horizon - self.window + 1
            if time_len < 1:
                raise ValueError(f"trajectory {i} has {traj_len} steps: too short for window={window} + horizon={horizon}")
            self._base.append(frame)
            self._first_example.append(ex)
            self._examples.append(time_len)
            frame, ex = frame + traj_len, ex + time_len
            feats.append(torch.from_numpy(f))
            conds.append(torch.as_tensor(np.asarray(_get(tr, "condition"), dtype=np.float32)))
            fm = _get(tr, "fixed_mask")
            if fm is not None:
                masks.append(torch.as_tensor(np.asarray(fm)).to(torch.float32).reshape(f.shape[1:]))
            vx = _get(tr, "vertices")
            if physical_system == "navier-stokes" and vx is not None and len(vx) > 0:
                verts.append(torch.as_tensor(np.asarray(vx, dtype=np.float32)))
            tt = _get(tr, "t")
            if tt is not None:
                t_all.append(torch.as_tensor(np.asarray(tt, dtype=np.float32)))
            self._meta.append(meta)
        self.n_examples, self.n_frames = ex, frame
        dev = self.device
        self.frames = torch.cat(feats, dim=0).contiguous().to(dev)              # (n_frames, C, H, W), uploaded once
        self.conditions = torch.stack(conds).contiguous().to(dev)               # (n_traj, Cs, H, W)
        self.fixed_masks = torch.stack(masks).contiguous().to(dev) if len(masks) == n else None
        self.vertices = torch.stack(verts).contiguous().to(dev) if len(verts) == n and n > 0 else None
        self.t = t_all if len(t_all) == n else None
        self._scalars = {}
        for key in ("time_step_size", "in_velocity"):
            if n > 0 and all(key in m for m in self._meta):
                self._scalars[key] = torch.tensor([float(m[key]) for m in self._meta], dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ index arithmetic (host)
    def __len__(self) -> int:
        return self.n_examples

    def origin(self, index: int) -> Tuple[int, int]:
        """global example index -> (trajectory, offset): trajectory-major like the reference's concatenation (:230-236)."""
        if index < 0:
            index += self.n_examples
        if not 0 <= index < self.n_examples:
            raise IndexError(f"example {index} out of range for {self.n_examples} examples")
        i = bisect.bisect_right(self._first_example, index) - 1
        return i, index - self._first_example[i]

    def first_frames(self, indices: Sequence[int]) -> Tuple[List[int], List[int]]:
        """-> (start frame in the store, trajectory index) per example."""
        trajs, first = [], []
        for g in indices:
            i, e = self.origin(int(g))
            trajs.append(i)
            first.append(self._base[i] + e)
        return first, trajs

    # ------------------------------------------------------------------ batches (device)
    def get_batch(self, indices: Sequence[int]) -> Dict[str, Any]:
        """The dict a reference DataLoader batch holds (`MyTensorDataset.__getitem__` + default collate), on the device."""
        first, trajs = self.first_frames(indices)
        batch: Dict[str, Any] = {"dynamics": window_gather(self.frames, first, self.frames_per_example)}
        batch["condition"] = window_gather(self.conditions, trajs, 1).squeeze(1)
        meta: Dict[str, Any] = {}
        if self.fixed_masks is not None:
            meta["fixed_mask"] = window_gather(self.fixed_masks, trajs, 1).squeeze(1) != 0
        if self.physical_system == "navier-stokes" and self.vertices is not None:
            meta["vertices"] = window_gather(self.vertices, trajs, 1).squeeze(1)
        if self.physical_system == "spring-mesh":
            meta["features"] = window_gather(self.frames, [self._base[i] for i in trajs], 1)  # first frame of the trajectory
        idx = torch.tensor(trajs, dtype=torch.long, device=self.device)
        for key, v in self._scalars.items():
            meta[key] = v.index_select(0, idx)
        if self.t is not None:  # (B, T) like the reference's collate; trajectories of unequal length (which the reference
            same = len({int(self.t[i].numel()) for i in set(trajs)}) == 1  # cannot collate) keep the column that is read: t[:, :1]
            meta["t"] = torch.stack([self.t[i] if same else self.t[i][:1] for i in trajs]).to(self.device)
        batch["metadata"] = meta
        return batch

    def batches(self, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                generator: Optional[torch.Generator] = None) -> Iterator[Dict[str, Any]]:
        """DataLoader replacement: sequential or shuffled example order (torch.randperm on the host, like RandomSampler)."""
        order = torch.randperm(self.n_examples, generator=generator).tolist() if shuffle else list(range(self.n_examples))
        for b in range(0, self.n_examples, batch_size):
            chunk = order[b:b + batch_size]
            if drop_last and len(chunk) < batch_size:
                return
            yield self.get_batch(chunk)

    @staticmethod
    def boundary_condition_kwargs(batch: Mapping[str, Any]) -> Dict[str, Any]:
        """`get_boundary_condition_kwargs` (physical_systems_benchmark.py:299-303)."""
        metadata = batch["metadata"]
        return dict(t0=metadata["t"][:, 0], dt=metadata["time_step_size"])
