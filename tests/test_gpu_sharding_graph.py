"""GPU tests of the round-2 sampler plumbing: dropout / noise streams keyed by the GLOBAL row (a row-sharded job equals the
un-sharded one bit for bit, ADVICE r1 "duplicated ensemble members"), CUDA-graph replay of `sample()` against plain
launches, stale-weight detection, experiment-level inference dropout reaching the forecaster."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _sample(dyf, ic, static, seed, **kw):
    torch.manual_seed(seed)
    dyf._calls = 0
    with torch.no_grad():
        return dyf.sample(ic, static_condition=static, **kw)


@pytest.mark.parametrize("dataset,horizon,rows", [("spring", 5, 8), ("sst", 3, 6)])
def test_sharded_rows_draw_the_masks_of_the_unsharded_job(dataset, horizon, rows):
    """dropout ON (and, for SST, "data+noise" noise): shards run with `row_offset` reproduce the rows of the full batch bit
    for bit, whatever the split; without the offset the second shard repeats the first shard's masks (the round-1 bug)."""
    from tests.gpu_helpers import build_dyffusion
    over = dict(additional_interpolation_steps=2) if dataset == "sst" else {}
    dyf = build_dyffusion(dataset, horizon=horizon, **over)
    ic, static = H.sampler_case_inputs("shard", dataset, rows)
    ic = ic[:1].repeat(rows, 1, 1, 1).cuda()  # identical inputs: only the random streams tell the members apart
    static = None if static is None else static[:1].repeat(rows, 1, 1, 1).cuda()
    full = _sample(dyf, ic, static, 7)
    k_last = sorted(full)[-1]
    assert not torch.equal(full[k_last][0], full[k_last][rows // 2])  # members differ
    for split in (rows // 2, 1, rows - 1):
        parts = []
        for b, e in ((0, split), (split, rows)):
            parts.append(_sample(dyf, ic[b:e].contiguous(), None if static is None else static[b:e].contiguous(), 7,
                                 row_offset=b))
        for k in full:
            assert torch.equal(torch.cat([parts[0][k], parts[1][k]]), full[k]), (split, k)
    # the hazard the offset removes: a shard that does not know its offset repeats rows 0.. of the job
    lo = _sample(dyf, ic[:rows // 2].contiguous(), None if static is None else static[:rows // 2].contiguous(), 7)
    hi = _sample(dyf, ic[rows // 2:].contiguous(), None if static is None else static[rows // 2:].contiguous(), 7)
    assert torch.equal(lo[k_last], hi[k_last])


@pytest.mark.parametrize("dataset,horizon,rows", [("spring", 6, 5), ("ns", 3, 2)])
def test_cuda_graph_replay_equals_plain_launches(dataset, horizon, rows):
    from tests.gpu_helpers import build_dyffusion
    import dyffusion_b200.engine as E
    plain = build_dyffusion(dataset, horizon=horizon, cuda_graph=False)
    graph = build_dyffusion(dataset, horizon=horizon, cuda_graph=True)
    ic, static = H.sampler_case_inputs("graph", dataset, rows)
    ic, static = ic.cuda(), static.cuda()
    outs = []
    for seed in (3, 4, 5, 6):  # run 1 plain (warms the caches), run 2 captures, runs 3-4 replay -- each with a new seed
        a = _sample(plain, ic, static, seed)
        n0 = E.launch_count()
        b = _sample(graph, ic.clone(), static.clone(), seed)  # fresh input tensors: the graph reads staged copies
        assert E.launch_count() > n0
        for k in a:
            assert torch.equal(a[k], b[k]), (seed, k)
        outs.append(b)
    assert not torch.equal(outs[2]["t1_preds"], outs[3]["t1_preds"])  # replays draw new dropout masks (seed in device memory)
    sampler = next(iter(graph._native_cache.values()))
    assert sampler.cuda_graph
    # runs 2-4 really went through cudaGraphLaunch (on PyTorch's default stream too: the legacy stream cannot be captured, the
    # sampler then runs on its own stream fenced with events) and none of the plain sampler's did
    assert sampler.graph_replays() == 3
    assert next(iter(plain._native_cache.values())).graph_replays() == 0


def test_graph_mode_on_a_side_stream_and_on_the_default_stream_agree():
    from tests.gpu_helpers import build_dyffusion
    dyf = build_dyffusion("spring", horizon=5, cuda_graph=True)
    ic, static = H.sampler_case_inputs("graph2", "spring", 4)
    ic, static = ic.cuda(), static.cuda()
    ref = [_sample(dyf, ic, static, seed) for seed in (1, 2, 3)]
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        out = [_sample(dyf, ic, static, seed) for seed in (1, 2, 3)]
    side.synchronize()
    for a, b in zip(ref, out):
        for k in a:
            assert torch.equal(a[k], b[k])


def test_weights_changed_in_place_are_repacked():
    """optimizer-style in-place updates bump the version counter; the reference's EMA swap uses `.data.copy_()`, which does
    not -- its LitEma object is hooked through the `ema_scope` assignment (src/experiment_types/_base_experiment.py:105-107)."""
    from contextlib import contextmanager

    from tests.gpu_helpers import build_backbone
    net = build_backbone("spring", "F", seed=1)
    x, cond = H.forward_inputs("spring", "F", rows=2)
    x, cond, t = x.cuda(), cond.cuda(), torch.tensor([0.0, 3.0]).cuda()
    with torch.no_grad():
        y0 = net(x, time=t, condition=cond)
        net.head.weight.mul_(0.5)          # what an optimizer step / nn.init does: tracked by `_version`
        y1 = net(x, time=t, condition=cond)
    assert not torch.equal(y0, y1) and H.rel_l2(y1, 0.5 * (y0 - net.head.bias.view(1, -1, 1, 1)) + net.head.bias.view(1, -1, 1, 1)) < 1e-2

    class Ema:  # the two LitEma methods that write `.data` (src/models/modules/ema.py:48-78)
        def __init__(self, model):
            self.shadow = [p.detach().clone() * 2.0 for p in model.parameters()]

        def store(self, params):
            self.saved = [p.detach().clone() for p in params]

        def copy_to(self, model):
            for p, s in zip(model.parameters(), self.shadow):
                p.data.copy_(s.data)

        def restore(self, params):
            for p, s in zip(params, self.saved):
                p.data.copy_(s.data)

    class Experiment:
        def __init__(self, model):
            self.model, self.model_ema = model, Ema(model)
            self.model.ema_scope = self.ema_scope

        @contextmanager
        def ema_scope(self):
            self.model_ema.store(self.model.parameters())
            self.model_ema.copy_to(self.model)
            try:
                yield None
            finally:
                self.model_ema.restore(self.model.parameters())

    exp = Experiment(net)
    with torch.no_grad():
        with exp.ema_scope():
            y_ema = net(x, time=t, condition=cond)
        y_back = net(x, time=t, condition=cond)
    assert not torch.equal(y_ema, y1)   # sampled with the EMA weights ...
    assert torch.equal(y_back, y1)      # ... and with the training weights again afterwards


def test_experiment_level_inference_dropout_reaches_the_forecaster():
    from tests.gpu_helpers import build_dyffusion
    dyf = build_dyffusion("spring", horizon=4, enable_interpolator_dropout=False, refine_intermediate_predictions=False,
                          sampling_type="naive")
    ic, static = H.sampler_case_inputs("fdrop", "spring", 3)
    ic, static = ic.cuda(), static.cuda()
    a = _sample(dyf, ic, static, 1)
    b = _sample(dyf, ic, static, 2)
    assert all(torch.equal(a[k], b[k]) for k in a)  # every dropout off: deterministic
    with dyf.inference_dropout_scope(True):
        c = _sample(dyf, ic, static, 1)
        d = _sample(dyf, ic, static, 2)
    assert not torch.equal(c["t4_preds"], a["t4_preds"]) and not torch.equal(c["t4_preds"], d["t4_preds"])
    e = _sample(dyf, ic, static, 1)
    assert all(torch.equal(a[k], e[k]) for k in a)


def test_clipping_survives_loading_a_torch_adamw_state_dict():
    from dyffusion_b200.optim import AdamW
    w = torch.randn(257, 3)
    ref_p = torch.nn.Parameter(w.clone().cuda())
    ref = torch.optim.AdamW([ref_p], lr=1e-2, betas=(0.9, 0.99), weight_decay=1e-4)
    ref_p.grad = torch.ones_like(ref_p)
    ref.step()
    p = torch.nn.Parameter(w.clone().cuda())
    opt = AdamW([p], lr=1e-2, betas=(0.9, 0.99), weight_decay=1e-4, max_grad_norm=1.0)
    opt.load_state_dict(ref.state_dict())
    assert opt.param_groups[0]["max_grad_norm"] == 1.0
    with torch.no_grad():
        p.copy_(ref_p)
    g = 100.0 * torch.randn_like(p)
    p.grad.copy_(g)
    opt.step()
    ref_p.grad = g.clone()
    torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
    ref.step()
    assert torch.allclose(p.detach(), ref_p.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dataset,rows,repeats", [("sst", 5, 12), ("ns", 2, 8), ("spring", 9, 12)])
def test_repeated_runs_are_bit_identical(dataset, rows, repeats):
    """Kernels launched with programmatic dependent launch may start before their predecessor has finished (they wait on
    `griddepcontrol.wait` before touching its output): a misplaced wait shows up as run-to-run differences.  Same seed, same
    inputs -> the same bits every time, for the plain-launch forward and for the graph-replayed sampler."""
    from tests.gpu_helpers import build_backbone, build_dyffusion
    net = build_backbone(dataset, "I", seed=2)
    x, c = H.forward_inputs(dataset, "I", rows=rows)
    t = torch.linspace(0.5, 2.5, rows).cuda()
    kw = {} if c is None else {"condition": c.cuda()}
    with torch.no_grad(), net.inference_dropout_scope(True):
        ref = None
        for _ in range(repeats):
            net._drop_stream = 0  # the same dropout stream every time
            y = net(x.cuda(), time=t, **kw)
            ref = y.clone() if ref is None else ref
            assert torch.equal(y, ref)
    dyf = build_dyffusion(dataset, horizon=3, cuda_graph=True)
    ic, static = H.sampler_case_inputs("rep", dataset, rows)
    first = None
    for _ in range(5):
        out = _sample(dyf, ic.cuda(), None if static is None else static.cuda(), 11)
        first = out if first is None else first
        for k in out:
            assert torch.equal(out[k], first[k])
