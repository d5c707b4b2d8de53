"""Row-sharded sampling across the GPUs of one node.

Rows (batch x ensemble members, member-major as produced by the reference's `get_ensemble_inputs`,
src/experiment_types/_base_experiment.py:503-538) are independent through the whole sampling loop, so they are split
contiguously over the ranks -- one process per GPU, no data-path collective -- and the per-rank forecasts are
exchanged with ONE all-gather at the end.  Weights are replicated; the dropout / noise streams are keyed by the global row
index (`row_offset`), so the sharded job equals the un-sharded one bit for bit and no ensemble member is drawn twice."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) row ranges, ceil(rows / world) per rank (trailing ranks may be short or empty)."""
    per = -(-rows // world)
    return [(min(r * per, rows), min((r + 1) * per, rows)) for r in range(world)]


def gather_rows(local: torch.Tensor, rows_total: int, group=None) -> torch.Tensor:
    """local: [K, rows_local, ...] on every rank (row shards in rank order) -> [K, rows_total, ...] on every rank."""
    world = dist.get_world_size(group)
    per = -(-rows_total // world)
    k, r_local = local.shape[0], local.shape[1]
    if r_local < per:  # short / empty tail shard: pad so that the collective is regular
        pad = local.new_zeros((k, per - r_local, *local.shape[2:]))
        local = torch.cat([local, pad], dim=1)
    local = local.contiguous()
    if dist.get_backend(group) == "nccl":
        out = local.new_empty((world, *local.shape))
        dist.all_gather_into_tensor(out, local, group=group)
    else:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
        out = torch.stack(parts)
    # [world, K, per, ...] -> [K, world * per, ...] keeps the global (member-major) row order
    out = out.transpose(0, 1).reshape(k, world * per, *local.shape[2:])
    return out[:, :rows_total]


def _row_offset_kw(diffusion, first_row: int) -> dict:
    """The engine drop-in keys its dropout / noise streams by the GLOBAL row (`sample_loop(row_offset=...)`): with the same
    torch seed on every rank (Lightning's seed_everything) the sharded job draws what the un-sharded job would.  Foreign
    diffusion modules (the reference's own) have no such argument: their ranks must be seeded differently by the caller."""
    import inspect

    fn = getattr(diffusion, "sample_loop", None)
    try:
        return {"row_offset": first_row} if fn is not None and "row_offset" in inspect.signature(fn).parameters else {}
    except (TypeError, ValueError):
        return {}


def _stacked(out: Dict[str, torch.Tensor], keys: List[str]) -> torch.Tensor:
    """[K, rows, ...] of the forecasts: the engine's sampler returns them as views of one such tensor -- use it as it is."""
    first = out[keys[0]]
    base = first._base
    if (base is not None and base.dim() == first.dim() + 1 and base.shape[0] == len(keys) and base.is_contiguous() and
            all(out[k].data_ptr() == base[i].data_ptr() and out[k].shape == base[i].shape for i, k in enumerate(keys))):
        return base
    return torch.stack([out[k] for k in keys])


def sample_sharded(diffusion, initial_condition: torch.Tensor, static_condition: Optional[torch.Tensor] = None,
                   group=None, **kwargs) -> Dict[str, torch.Tensor]:
    """`diffusion.sample` over this rank's shard of the rows + one all-gather; every rank returns the full dict.
    `initial_condition` / `static_condition` hold ALL rows on every rank (as Lightning's replicated eval batch would)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return diffusion.sample(initial_condition, static_condition=static_condition, **kwargs)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rows = initial_condition.shape[0]
    b, e = shard_bounds(rows, world)[rank]
    keys: List[str]
    if e > b:
        out = diffusion.sample(initial_condition[b:e],
                               static_condition=None if static_condition is None else static_condition[b:e],
                               **_row_offset_kw(diffusion, b), **kwargs)
        keys = list(out.keys())
        local = _stacked(out, keys)
    else:  # more ranks than rows: take the output structure from a one-row dry description
        keys, local = [], None
    # all ranks must agree on the keys; rank 0 always owns rows
    obj = [keys]
    dist.broadcast_object_list(obj, src=0, group=group)
    keys = obj[0]
    if local is None:
        c = getattr(diffusion, "num_input_channels", initial_condition.shape[1])
        local = initial_condition.new_zeros((len(keys), 0, c, *initial_condition.shape[2:]))
    full = gather_rows(local, rows, group=group)
    return {k: full[i] for i, k in enumerate(keys)}


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Data-parallel gradient synchronisation for the training tier: Lightning's DDP averages the
    gradients in ~25 MB buckets while the backward runs (`trainer=ddp`); with `dyffusion_b200.optim.AdamW` every gradient is a
    view of ONE flat arena, so the whole exchange is a single in-place all-reduce of that arena (NCCL picks NVLS / ring over
    NVSwitch; one launch, no bucketing logic) followed by the 1/world scale.  No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:  # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat
