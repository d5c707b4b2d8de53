"""CPU-only tests of the boundary: the C-ABI library loads and exports every symbol the header declares, the
drop-in classes reproduce the reference's state-dict keys / schedule logic / error behaviour, and nothing in the
product path falls back to CPU arithmetic or imports the oracle."""
import json
import os
import re
import subprocess

import pytest
import torch

from oracle import configs as C
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = H.golden_json("state_shapes.json")
KAT = H.golden_json("schedule_kat.json")


def test_library_exports_every_declared_symbol():
    from dyffusion_b200 import engine as E

    header = open(os.path.join(ROOT, "include", "dyffusion_b200.h")).read()
    declared = set(re.findall(r"\b(dyf_[a-z_0-9]+)\s*\(", header))
    assert declared == set(E.EXPORTED), declared ^ set(E.EXPORTED)
    nm = subprocess.run(["nm", "-D", "--defined-only", E.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dyf_[a-z_0-9]+)", nm))
    assert declared <= exported, declared - exported
    assert E.LIB.dyf_abi_version() == E.ABI_VERSION == 2
    assert E.launch_count() >= 0


def _build(dataset, role):
    from tests.gpu_helpers import build_backbone
    return build_backbone(dataset, role, device="cpu")


@pytest.mark.parametrize("dataset,role", [("ns", "F"), ("ns", "I"), ("spring", "F"), ("spring", "I"), ("sst", "F"),
                                          ("sst", "I")])
def test_state_dict_contract(dataset, role):
    """Weight-layout contract: keys and shapes equal the reference's, so its checkpoints load."""
    m = _build(dataset, role)
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == SHAPES[f"{dataset}_{role}"]
    m.load_state_dict({k: torch.zeros(s) if not k.endswith("num_batches_tracked") else torch.zeros((), dtype=torch.long)
                       for k, s in SHAPES[f"{dataset}_{role}"].items()}, strict=True)
    for attr in ("num_input_channels", "num_output_channels", "num_conditional_channels", "spatial_shape", "criterion",
                 "hparams", "num_params", "ema_scope"):
        assert hasattr(m, attr)
    assert m.num_params == sum(int(torch.tensor(s).prod()) if s else 1 for k, s in SHAPES[f"{dataset}_{role}"].items()
                               if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU rather than compute on the host."""
    from dyffusion_b200.engine import EngineError

    m = _build("spring", "F").eval()
    x = torch.zeros(1, 4, 10, 10)
    with torch.no_grad(), pytest.raises((EngineError, RuntimeError)):
        m(x, time=torch.zeros(1), condition=torch.zeros(1, 5, 10, 10))


def test_no_cpu_fallback_for_the_widening_rows():
    """Ensemble metrics and boundary conditions: CPU tensors are refused, and the C entry points themselves fail with
    DYF_ERR_CUDA when there is no device (no host arithmetic anywhere behind the ABI)."""
    import ctypes as C

    import dyffusion_b200.engine as E
    from dyffusion_b200.boundary import boundary_conditions
    from dyffusion_b200.metrics import evaluate_ensemble_prediction

    with pytest.raises(E.EngineError):
        evaluate_ensemble_prediction(torch.zeros(3, 2, 4, 5, 5), torch.zeros(2, 4, 5, 5))
    with pytest.raises(E.EngineError):
        boundary_conditions("spring-mesh", torch.zeros(2, 4, 10, 10), torch.zeros(2, 4, 10, 10), {})
    if not torch.cuda.is_available():
        buf = (C.c_double * 8)()
        fbuf = (C.c_float * 64)()
        ws = (C.c_char * 4096)()
        rc = E.LIB.dyf_ensemble_metrics(C.addressof(fbuf), C.addressof(fbuf), 2, 1, 4, C.addressof(buf), None, C.addressof(ws), 4096, None)
        assert rc == -2 and b"no CPU fallback" in E.LIB.dyf_last_error()
        rc = E.LIB.dyf_boundary_conditions_spring_mesh(C.addressof(fbuf), C.addressof(ws), C.addressof(fbuf), 1, 1, 2, 2, None)
        assert rc == -2 and b"no CPU fallback" in E.LIB.dyf_last_error()


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dyffusion_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "oracle/" in txt:
                    bad.append(f)
    assert not bad, bad


def test_schedule_logic_matches_reference_kat():
    from dyffusion_b200.diffusion.schedule import DiffusionSchedule

    for rec in KAT["schedules"]:
        a = rec["args"]
        mk = lambda: DiffusionSchedule(a["timesteps"], a.get("schedule", "before_t1_only"),
                                       a.get("additional_interpolation_steps", 0),
                                       a.get("additional_interpolation_steps_factor", 0),
                                       a.get("interpolate_before_t1", True))
        if "error" in rec:
            with pytest.raises((AssertionError, ValueError, IndexError)):
                mk().parse_sampling_schedule(rec["spec"])
            continue
        s = mk()
        sched = s.parse_sampling_schedule(rec["spec"])
        assert s.num_timesteps == rec["num_timesteps"]
        assert [float(v) for v in sched] == rec["schedule"]
        assert all(isinstance(v, int) for v in sched) == rec["all_int"]
        assert [float(s.interpolation_time(d)) for d in range(s.num_timesteps)] == rec["tau"]
        assert {str(k): float(v) for k, v in s.dynamical_steps.items()} == rec["dynamical"]


def test_dyffusion_dropin_surface_and_errors():
    from tests.gpu_helpers import build_dyffusion
    from dyffusion_b200.diffusion.dyffusion import DYffusion

    dyf = build_dyffusion("ns", device="cpu", horizon=16)
    assert dyf.num_timesteps == 16 and dyf.sampling_schedule == list(range(16))
    assert dyf.dynamical_steps == {d: d for d in range(1, 16)}
    assert "dyffusion" in (type(dyf).__module__ + "." + type(dyf).__name__).lower()
    assert "static_condition" in __import__("inspect").signature(dyf.p_losses).parameters
    assert "num_predictions" in __import__("inspect").signature(dyf.sample_loop).parameters
    assert dyf.interpolator.hparams.num_predictions == 1 and dyf.interpolator_horizon == 16
    assert all(not p.requires_grad for p in dyf.interpolator.parameters())
    assert float(dyf.diffusion_step_to_interpolation_step(torch.tensor(3.0))) == 3.0
    sst = build_dyffusion("sst", device="cpu", horizon=7, backbone_arch_override="spring")
    assert sst.num_timesteps == 32 and sst.dynamical_steps == {26: 1, 27: 2, 28: 3, 29: 4, 30: 5, 31: 6}
    sst.sampling_schedule = "every5th"
    assert sst.sampling_schedule == [0, 1, 6, 11, 16, 21, 26, 27, 28, 29, 30, 31]
    with pytest.raises(ValueError):
        sst.sampling_schedule = "bogus"
    with pytest.raises(ValueError):  # interpolator horizon mismatch (dyffusion.py:472-478)
        build_dyffusion("ns", device="cpu", horizon=16, interpolator_horizon=8)
    with pytest.raises(ValueError):
        DYffusion(model=None, timesteps=4)
    with pytest.raises(ValueError):
        build_dyffusion("ns", device="cpu", horizon=4, forward_conditioning="nonsense")
