"""GPU parity tests proper: the CUDA path (through the C ABI / the drop-in classes) against the oracle on the
same seeded inputs and against the committed reference-generated golden vectors.

Stated tolerances (16-bit operands and activations between layers -- fp16 by default --, fp32 accumulate / epilogue /
statistics), the same for all three backbones: rel-L2 <= 1e-2 per forward, <= 5e-2 per chained sampler
trajectory.  (Round 1 stored bf16 and needed 2e-2 / 1e-1 for the four-times-deeper SST `Unet`; measured now: 1.4e-3 per
SST forward, 8.3e-3 over the 93-call SST trajectory, see tests/test_gpu_fullhorizon.py.)
Host-side bookkeeping (keys, call structure) must be exact."""
import pytest
import torch

from oracle import configs as C
from oracle import dyffusion_oracle as O
from oracle.synth import synth_state_dict, synth_tensor
from tests import helpers as H

pytestmark = pytest.mark.gpu
FWD_TOLS = {"ns": 1e-2, "spring": 1e-2, "sst": 1e-2}
TRAJ_TOLS = {"ns": 5e-2, "spring": 5e-2, "sst": 5e-2}
SHAPES = H.golden_json("state_shapes.json")
KAT = H.golden_json("schedule_kat.json")
BUILT = [("ns", "F"), ("ns", "I"), ("spring", "F"), ("spring", "I"), ("sst", "F"), ("sst", "I")]


def _engine_loaded():
    import dyffusion_b200.engine as E
    assert E.LIB is not None


@pytest.mark.parametrize("dataset,role", BUILT)
def test_forward_vs_oracle_and_golden(dataset, role):
    from tests.gpu_helpers import build_backbone
    _engine_loaded()
    tag = f"{dataset}_{role}"
    g = H.golden_pt(f"fwd_{tag}.pt")
    net = build_backbone(dataset, role, seed=g["weight_seed"])
    x, cond = H.forward_inputs(dataset, role, rows=g["rows"])
    with torch.no_grad():
        y = net(x.cuda(), time=g["time"].cuda(), condition=None if cond is None else cond.cuda()).cpu()
        sd = synth_state_dict(SHAPES[tag], seed=g["weight_seed"])
        y_or = H.oracle_net(dataset, role, sd)(x, g["time"], cond)
    assert torch.isfinite(y).all()
    print(f"{tag}: rel-L2 vs oracle {H.rel_l2(y, y_or):.2e}, vs reference golden {H.rel_l2(y, g['y']):.2e}")
    assert H.rel_l2(y, y_or) <= FWD_TOLS[dataset], H.rel_l2(y, y_or)
    assert H.rel_l2(y, g["y"]) <= FWD_TOLS[dataset], H.rel_l2(y, g["y"])


@pytest.mark.parametrize("rows", [1, 3, 5])
def test_forward_ragged_rows_and_row_independence(rows):
    """Rows are independent (the shard axis): a batch equals its rows run one by one, bit for bit."""
    from tests.gpu_helpers import build_backbone
    net = build_backbone("spring", "I", seed=1)
    x = synth_tensor("rag.x", (rows, 8, 10, 10)).cuda()
    c = synth_tensor("rag.c", (rows, 1, 10, 10), kind="mask").cuda()
    t = torch.linspace(0.5, 3.0, rows).cuda()
    with torch.no_grad():
        y = net(x, time=t, condition=c)
        y1 = torch.cat([net(x[i:i + 1], time=t[i:i + 1], condition=c[i:i + 1]) for i in range(rows)])
    assert torch.equal(y, y1)


def test_ns_row_independence_bitexact():
    from tests.gpu_helpers import build_backbone
    net = build_backbone("ns", "F", seed=1)
    x = synth_tensor("nsr.x", (3, 3, 221, 42)).cuda()
    c = synth_tensor("nsr.c", (3, 2, 221, 42), kind="mask").cuda()
    t = torch.tensor([0.0, 4.0, 9.0]).cuda()
    with torch.no_grad():
        y = net(x, time=t, condition=c)
        y2 = net(x[2:3], time=t[2:3], condition=c[2:3])
    assert torch.equal(y[2:3], y2)


def test_dropout_statistics_and_determinism():
    """In-kernel Philox dropout cannot match torch's stream bit for bit: check determinism under a
    fixed seed, independence across calls, and that the ensemble mean approaches the dropout-free output."""
    from tests.gpu_helpers import build_backbone
    net = build_backbone("spring", "I", seed=1)
    x = synth_tensor("dr.x", (1, 8, 10, 10)).cuda().repeat(512, 1, 1, 1)
    c = synth_tensor("dr.c", (1, 1, 10, 10), kind="mask").cuda().repeat(512, 1, 1, 1)
    t = torch.full((512,), 2.0).cuda()
    with torch.no_grad():
        y0 = net(x, time=t, condition=c)
        torch.manual_seed(123)
        net._drop_stream = 0
        with net.inference_dropout_scope(True):
            ya = net(x, time=t, condition=c)
            yb = net(x, time=t, condition=c)
        torch.manual_seed(123)
        net._drop_stream = 0
        with net.inference_dropout_scope(True):
            ya2 = net(x, time=t, condition=c)
    assert torch.equal(ya, ya2)            # deterministic given (seed, call index)
    assert not torch.equal(ya, yb)         # a new call draws new masks
    assert not torch.equal(ya[0], ya[1])   # rows draw independent masks
    assert H.rel_l2(ya.mean(0), y0[0]) < 0.1  # p = 0.05: ensemble mean ~ deterministic output
    assert (ya.std(0) > 0).float().mean() > 0.9


def test_dropout_mask_rate_and_replay():
    import dyffusion_b200.engine as E
    for p in (0.05, 0.15, 0.6):
        m = E.debug_dropout_mask(7, 3, 2, p, 1 << 20).float()
        assert abs(float(m.mean()) - (1 - p)) < 3e-3
        assert torch.equal(m, E.debug_dropout_mask(7, 3, 2, p, 1 << 20).float())
        assert not torch.equal(m, E.debug_dropout_mask(7, 4, 2, p, 1 << 20).float())


SAMPLERS = [k for k in KAT if k != "schedules" and KAT[k]["overrides"].get("forward_conditioning", C.DIFFUSION[KAT[k]["dataset"]]["forward_conditioning"]) != "data+noise"]


@pytest.mark.parametrize("name", SAMPLERS)
def test_sampler_vs_oracle_and_golden(name):
    from tests.gpu_helpers import build_dyffusion
    meta = KAT[name]
    ds = meta["dataset"]
    g = H.golden_pt(f"sample_{name}.pt")
    dyf = build_dyffusion(ds, enable_interpolator_dropout=False, **meta["overrides"])
    assert [float(s) for s in dyf.sampling_schedule] == meta["sampling_schedule"]
    ic, static = H.sampler_case_inputs(name, ds, g["rows"])
    with torch.no_grad():
        out = dyf.predict_forward(ic.cuda(), condition=None if static is None else static.cuda())
        out_py = dyf._sample_loop_python(ic.cuda(), None if static is None else static.cuda(), None, None)[1]
    assert sorted(out) == meta["keys"] == sorted(out_py)
    for k, v in g["preds"].items():
        assert H.rel_l2(out[k].cpu(), v) <= TRAJ_TOLS[ds], (k, H.rel_l2(out[k].cpu(), v))
        # the native loop and the Python-driven loop launch the same kernels on the same data
        assert torch.equal(out[k], out_py[k]), k


def test_sampler_with_dropout_is_seeded_and_stochastic():
    from tests.gpu_helpers import build_dyffusion
    dyf = build_dyffusion("spring", horizon=5)
    ic, static = H.sampler_case_inputs("sd", "spring", 4)
    ic, static = ic.cuda(), static.cuda()
    with torch.no_grad():
        torch.manual_seed(5); dyf._calls = 0
        a = dyf.sample(ic, static_condition=static)
        b = dyf.sample(ic, static_condition=static)
        torch.manual_seed(5); dyf._calls = 0
        a2 = dyf.sample(ic, static_condition=static)
    assert sorted(a) == [f"t{i}_preds" for i in range(1, 6)]
    assert all(torch.equal(a[k], a2[k]) for k in a)
    assert not torch.equal(a["t1_preds"], b["t1_preds"])


def test_error_behaviour_on_device():
    from tests.gpu_helpers import build_backbone
    net = build_backbone("spring", "F", seed=1)
    x = torch.zeros(2, 4, 10, 10, device="cuda")
    with torch.no_grad():
        with pytest.raises(ValueError):
            net(x, time=torch.zeros(2, device="cuda"), condition=torch.zeros(2, 4, 10, 10, device="cuda"))
        with pytest.raises(ValueError):
            net(torch.zeros(2, 4, 9, 10, device="cuda"), time=torch.zeros(2, device="cuda"),
                condition=torch.zeros(2, 5, 10, 10, device="cuda"))


def test_forward_is_bit_reproducible():
    """No float atomics anywhere on the path: repeated launches give bit-identical results (GroupNorm statistics are
    reduced in a fixed order)."""
    from tests.gpu_helpers import build_backbone
    net = build_backbone("ns", "I", seed=1)
    x, cond = H.forward_inputs("ns", "I", rows=3)
    t = torch.tensor([1.0, 2.0, 3.5]).cuda()
    with torch.no_grad():
        ys = [net(x.cuda(), time=t, condition=cond.cuda()) for _ in range(4)]
    assert all(torch.equal(ys[0], y) for y in ys[1:])


def test_sst_data_plus_noise_sampler():
    """forward_conditioning="data+noise" (SST): the Python-driven loop with the golden run's injected noise must match
    the reference golden; the native loop draws its own Philox noise, so it is checked for seeding / stochasticity."""
    from tests.gpu_helpers import build_dyffusion
    name = "sst_h3_k2"
    meta = KAT[name]
    g = H.golden_pt(f"sample_{name}.pt")
    dyf = build_dyffusion("sst", enable_interpolator_dropout=False, **meta["overrides"])
    ic, _ = H.sampler_case_inputs(name, "sst", g["rows"])
    ic = ic.cuda()
    calls = {"n": 0}

    def fake(t):
        calls["n"] += 1
        return synth_tensor(f"{name}.noise{calls['n'] - 1}", tuple(t.shape)).to(t.device)

    real = torch.randn_like
    torch.randn_like = fake
    try:
        with torch.no_grad():
            out_py = dyf._sample_loop_python(ic, None, None, None)[1]
    finally:
        torch.randn_like = real
    assert calls["n"] == meta["noise_draws"] and sorted(out_py) == meta["keys"]
    for k, v in g["preds"].items():
        assert H.rel_l2(out_py[k].cpu(), v) <= TRAJ_TOLS["sst"], (k, H.rel_l2(out_py[k].cpu(), v))
    with torch.no_grad():
        torch.manual_seed(11); dyf._calls = 0
        a = dyf.sample(ic)
        b = dyf.sample(ic)
        torch.manual_seed(11); dyf._calls = 0
        a2 = dyf.sample(ic)
    assert sorted(a) == meta["keys"] and all(torch.isfinite(v).all() for v in a.values())
    assert all(torch.equal(a[k], a2[k]) for k in a) and not torch.equal(a["t1_preds"], b["t1_preds"])
