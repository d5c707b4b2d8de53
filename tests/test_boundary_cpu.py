"""CPU tests of the boundary-condition oracle (oracle/boundary_oracle.py): against the reference method itself where
/root/reference is importable (the build container), plus self-contained known answers that run anywhere."""
import math
import types

import pytest
import torch

from oracle import boundary_oracle as B
from oracle import ref_shims


def _ns_case(b=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    preds = torch.randn(b, 3, 221, 42, generator=g)
    meta = {"fixed_mask": torch.rand(b, 3, 221, 42, generator=g) < 0.1,
            "vertices": torch.rand(b, 2, 221, 42, generator=g) * 0.41,
            "in_velocity": torch.rand(b, 1, generator=g) + 0.5}
    return preds, torch.zeros(b, 3, 221, 42), meta


def _ns_ensemble_case(members=3, b=2, seed=9):
    """(members, batch, 3, H, W) predictions, as the rollout hands them over (forecasting_multi_horizon.py:175-182)."""
    _, tg, meta = _ns_case(b=b, seed=seed)
    return torch.randn(members, b, 3, 221, 42, generator=torch.Generator().manual_seed(1)), tg, meta


def _spring_case(b=4, lead=None, seed=1):
    g = torch.Generator().manual_seed(seed)
    shape = (b, 4, 10, 10) if lead is None else (lead, b, 4, 10, 10)
    preds = torch.randn(*shape, generator=g)
    meta = {"fixed_mask": torch.rand(b, 4, 10, 10, generator=g) < 0.3, "features": torch.randn(b, 6, 4, 10, 10, generator=g)}
    return preds, torch.zeros(b, 4, 10, 10), meta


def test_known_answers():
    preds, tg, meta = _ns_case()
    out = B.boundary_conditions("navier-stokes", preds.clone(), tg, meta, time=0.7)
    m = meta["fixed_mask"].clone()
    m[:, 0, 0, :] = False
    assert (out[m] == 0).all()
    y = meta["vertices"][1, 1, 0, :]
    want = float(meta["in_velocity"][1]) * 4 * y * (0.41 - y) / (0.41 * 0.41) * (1 - math.exp(-5 * 0.7))
    assert torch.allclose(out[1, 0, 0, :], want)
    keep = ~meta["fixed_mask"]
    keep[:, 0, 0, :] = False
    assert torch.equal(out[keep], preds[keep])
    preds, tg, meta = _spring_case(lead=3)
    out = B.boundary_conditions("spring-mesh", preds.clone(), tg, meta)
    fm = meta["fixed_mask"]
    assert torch.equal(out[:, ~fm], preds[:, ~fm])
    assert torch.equal(out[1, :, 2:][fm[:, 2:]], meta["features"][:, 0, 2:][fm[:, 2:]])
    assert (out[2, :, :2][fm[:, :2]] == 0).all()
    with pytest.raises(NotImplementedError):
        B.boundary_conditions("pendulum", preds, tg, meta)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference (build container only)")
def test_oracle_equals_reference_method():
    ref_shims.install()
    from src.datamodules.physical_systems_benchmark import PhysicalSystemsBenchmarkDataModule as DM
    for system, case, kw in (("navier-stokes", _ns_case(), dict(time=1.3)),
                             ("navier-stokes", _ns_case(seed=5), dict(time=torch.tensor([0.1, 0.5, 2.0]))),
                             ("navier-stokes", _ns_ensemble_case(), dict(time=torch.tensor([0.3, 1.1]))),
                             ("spring-mesh", _spring_case(), {}), ("spring-mesh", _spring_case(lead=5), {})):
        preds, tg, meta = case
        fake = types.SimpleNamespace(hparams=types.SimpleNamespace(physical_system=system))
        want = DM.boundary_conditions(fake, preds.clone(), tg, meta, **kw)
        got = B.boundary_conditions(system, preds.clone(), tg, meta, **kw)
        assert torch.equal(got, want), system
