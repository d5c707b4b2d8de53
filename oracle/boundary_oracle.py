"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of `PhysicalSystemsBenchmarkDataModule.boundary_conditions`
(`src/datamodules/physical_systems_benchmark.py:245-297`), the checker for `dyf_boundary_conditions_*`.
Only tests/ may import this.  Pinned by `tests/golden/make_golden.py`-style differential runs against the reference method
itself in the build container (tests/test_boundary_cpu.py imports the reference class through oracle/ref_shims.py when
/root/reference is present and compares; the restatement alone is used elsewhere)."""
from __future__ import annotations

import math

import torch


def boundary_conditions(physical_system: str, preds: torch.Tensor, targets: torch.Tensor, metadata, time=None) -> torch.Tensor:
    batch_size = targets.shape[0]
    if physical_system == "navier-stokes":  # :253-276
        for b_i in range(batch_size):
            t_i = time if isinstance(time, float) else time[b_i].item()
            in_velocity = float(metadata["in_velocity"][b_i].item())
            fixed = metadata["fixed_mask"][b_i, ...]
            assert fixed.shape == preds.shape[-3:]
            vertex_y = metadata["vertices"][b_i, 1, 0, :]
            left_idx = torch.zeros(tuple(preds.shape[-3:]), dtype=torch.bool)
            left_idx[0, 0, :] = True
            left = in_velocity * 4 * vertex_y * (0.41 - vertex_y) / (0.41 * 0.41) * (1 - math.exp(-5 * t_i))
            preds[b_i, ..., fixed] = 0
            preds[b_i, ..., left_idx] = left.unsqueeze(0)
    elif physical_system == "spring-mesh":  # :277-287
        for b_i in range(batch_size):
            fixed = metadata["fixed_mask"][b_i]
            assert fixed.shape[0] == 4
            base_q = metadata["features"][b_i, 0, 2:]
            bc = torch.cat([torch.zeros_like(base_q), base_q], dim=0)
            if preds.ndim == 5 and preds.shape[1] == batch_size:
                preds[:, b_i, ...] = torch.where(fixed, bc, preds[:, b_i, ...])
            else:
                preds[b_i, ...] = torch.where(fixed, bc, preds[b_i, ...])
    else:
        raise NotImplementedError(f"Boundary conditions for {physical_system} not implemented")
    return preds
