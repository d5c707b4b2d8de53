"""GPU parity of `dyf_boundary_conditions_*` (through dyffusion_b200.boundary -> ctypes -> C ABI) against the oracle
restatement of physical_systems_benchmark.py:245-297.  Masked writes are exact; the inflow profile is fp32 arithmetic in
the reference's operation order: tolerance 2 ulp (rel 3e-7)."""
import pytest
import torch

from oracle import boundary_oracle as B
from tests.test_boundary_cpu import _ns_case, _spring_case

pytestmark = pytest.mark.gpu


def _cuda_meta(meta):
    return {k: v.cuda() for k, v in meta.items()}


@pytest.mark.parametrize("time", [0.9, "per_sample"])
def test_navier_stokes(time):
    from dyffusion_b200.boundary import boundary_conditions
    preds, tg, meta = _ns_case(b=5, seed=3)
    t = torch.tensor([0.0, 0.2, 0.8, 1.5, 4.0]) if time == "per_sample" else time
    want = B.boundary_conditions("navier-stokes", preds.clone(), tg, meta, time=t)
    p = preds.clone().cuda()
    got = boundary_conditions("navier-stokes", p, tg.cuda(), _cuda_meta(meta), time=t)
    assert got.data_ptr() == p.data_ptr()  # in place, like the reference
    row0 = torch.zeros_like(want, dtype=torch.bool)
    row0[:, 0, 0, :] = True
    assert torch.equal(got.cpu()[~row0], want[~row0])
    assert torch.allclose(got.cpu()[row0], want[row0], rtol=3e-7, atol=1e-9)


@pytest.mark.parametrize("lead", [None, 6])
def test_spring_mesh(lead):
    from dyffusion_b200.boundary import boundary_conditions
    preds, tg, meta = _spring_case(b=7, lead=lead, seed=4)
    want = B.boundary_conditions("spring-mesh", preds.clone(), tg, meta)
    got = boundary_conditions("spring-mesh", preds.clone().cuda(), tg.cuda(), _cuda_meta(meta))
    assert torch.equal(got.cpu(), want)


def test_errors():
    from dyffusion_b200.boundary import boundary_conditions
    import dyffusion_b200.engine as E
    preds, tg, meta = _spring_case()
    with pytest.raises(NotImplementedError):
        boundary_conditions("pendulum", preds.cuda(), tg.cuda(), _cuda_meta(meta))
    with pytest.raises(E.EngineError):
        boundary_conditions("spring-mesh", preds, tg, meta)
