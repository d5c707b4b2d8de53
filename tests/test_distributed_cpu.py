"""world_size-2 gloo tests (CPU) of the N>1 host logic: contiguous row shards, one all-gather, global row order."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dyffusion_b200.distributed import allreduce_mean_, gather_rows, sample_sharded, shard_bounds


class _FakeDiffusion:
    """Stands in for the engine-backed sampler on CPU: output slot i of a row is a known function of that row."""
    num_input_channels = 3

    def sample(self, ic, static_condition=None, **kw):
        s = 0 if static_condition is None else static_condition.sum(dim=1, keepdim=True)
        return {f"t{i}_preds": ic[:, -3:] * i + s for i in (1, 2, 3)}


class _FakeEngineDiffusion(_FakeDiffusion):
    """Like the engine drop-in, its `sample_loop` takes `row_offset` (index of the shard's first row in the un-sharded job, what
    the dropout / noise streams are keyed by): the outputs record the GLOBAL row every local row believes it is."""

    def sample_loop(self, initial_condition, static_condition=None, row_offset=0):
        raise NotImplementedError

    def sample(self, ic, static_condition=None, row_offset=0, **kw):
        out = super().sample(ic, static_condition, **kw)
        glob = torch.arange(ic.shape[0], dtype=ic.dtype).view(-1, 1, 1, 1) + float(row_offset)
        out["t1_preds"] = out["t1_preds"] * 0 + glob
        return out


def _rollout_is_rank_invariant(rows):
    """The autoregressive rollout with its sampler calls sharded over the ranks returns, on every rank, what one process
    returns: member-major row order survives shard -> all-gather -> (N, B) un-stacking -> hand-off."""
    from dyffusion_b200.rollout import MultiHorizonRollout

    class Diff(_FakeDiffusion):
        num_timesteps = 3
        hparams = {"timesteps": 3}

        def sample_loop(self, initial_condition, static_condition=None, log_every_t=None, num_predictions=None):
            raise NotImplementedError

        def predict_forward(self, inputs, condition=None, metadata=None, **kw):
            return self.sample(inputs, static_condition=condition)

    members = 2
    g = torch.Generator().manual_seed(1)
    batch = {"dynamics": torch.randn(rows, 7, 3, 5, 4, generator=g), "condition": torch.randn(rows, 2, 5, 4, generator=g)}
    bc = lambda preds, targets, metadata, time: preds.mul_(0.5).add_(time)
    kw = dict(horizon=3, num_predictions=members, autoregressive_steps=1)
    want = MultiHorizonRollout(Diff(), **kw).evaluation_step(batch, "test", boundary_conditions=bc)
    got = MultiHorizonRollout(Diff(), group=dist.group.WORLD, **kw).evaluation_step(batch, "test", boundary_conditions=bc)
    return list(got) == list(want) and all(torch.equal(got[k], want[k]) for k in want) and \
        tuple(got["t6_preds"].shape) == (members, rows, 3, 5, 4)


def _worker(rank, world, port, rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        ic = torch.randn(rows, 6, 5, 4, generator=g)
        st = torch.randn(rows, 2, 5, 4, generator=g)
        out = sample_sharded(_FakeDiffusion(), ic, st)
        ref = _FakeDiffusion().sample(ic, st)
        ok = sorted(out) == sorted(ref) and all(torch.equal(out[k], ref[k]) for k in ref)
        # every rank hands its shard's first row to the sampler: gathered, the recorded global rows are 0 .. rows-1
        rec = sample_sharded(_FakeEngineDiffusion(), ic, st)["t1_preds"]
        ok = ok and torch.equal(rec[:, 0, 0, 0], torch.arange(rows, dtype=rec.dtype))
        # raw gather keeps rank order and drops the padding of a short tail shard
        b, e = shard_bounds(rows, world)[rank]
        local = torch.arange(b, e, dtype=torch.float32).view(1, -1, 1).repeat(2, 1, 3)
        full = gather_rows(local, rows)
        ok = ok and torch.equal(full[0, :, 0], torch.arange(rows, dtype=torch.float32)) and full.shape == (2, rows, 3)
        # gradient arena: one in-place all-reduce = the mean over ranks (DDP semantics)
        arena = torch.arange(10, dtype=torch.float32) * (rank + 1)
        allreduce_mean_(arena)
        ok = ok and torch.allclose(arena, torch.arange(10, dtype=torch.float32) * ((world + 1) / 2))
        ok = ok and _rollout_is_rank_invariant(rows)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _run_world(world, rows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * rows + 131 * world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert res == {r: True for r in range(world)}


@pytest.mark.parametrize("rows", [8, 5, 1])
def test_sharded_sampling_world2(rows):
    _run_world(2, rows)


@pytest.mark.parametrize("rows", [6, 3])
def test_sharded_sampling_world4_with_short_and_empty_shards(rows):
    """6 rows over 4 ranks -> shards of 2, 2, 2, 0; 3 rows -> 1, 1, 1, 0: the last rank owns no rows and still takes part in
    the all-gather (keys come from rank 0), every rank ends with all rows in the global order."""
    _run_world(4, rows)


def test_shard_bounds():
    assert shard_bounds(64, 8) == [(8 * r, 8 * r + 8) for r in range(8)]
    assert shard_bounds(300, 8)[0] == (0, 38) and shard_bounds(300, 8)[-1] == (266, 300)
    assert shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
