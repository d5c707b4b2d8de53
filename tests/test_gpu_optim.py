"""GPU parity of `dyffusion_b200.optim.AdamW` (-> ctypes -> `dyf_adamw_step`) against torch.optim.AdamW +
clip_grad_norm_ run on CPU by oracle/optim_oracle.py (the implementation the reference itself uses).
Stated tolerance: 1e-5 of each tensor's scale on parameters and moments, rel-L2 1e-4 on the accumulated update (fp32
arithmetic in torch's operation order; differences are FMA contraction / 1-ulp level)."""
import pytest
import torch

from oracle import optim_oracle as OO
from tests import helpers as H
from tests.test_optim_cpu import HYPER, case, close, update_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("max_norm,scale", [(None, 1.0), (1.0, 1.0), (1.0, 1e-4)])
def test_steps_equal_torch_adamw(max_norm, scale):
    from dyffusion_b200.optim import AdamW
    import dyffusion_b200.engine as E
    params, grads = case(steps=6, scale=scale)
    want_p, want_m, want_v, norms, _ = OO.reference_steps(params, grads, max_grad_norm=max_norm, **HYPER)
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = AdamW(ps, max_grad_norm=max_norm, **HYPER)
    before = E.launch_count()
    for k, gs in enumerate(grads):
        opt.zero_grad()
        for p, g in zip(ps, gs):
            if k % 2 == 0:
                p.grad.add_(g.cuda())       # accumulate into the arena view (what autograd does)
            else:
                p.grad = g.clone().cuda()   # a foreign gradient tensor: collected at step()
        opt.step()
        if max_norm:
            assert abs(float(opt.grad_norm()) - norms[k]) <= 2e-6 * norms[k]
    assert E.launch_count() - before == len(grads) * (3 if max_norm else 1)
    for i, p in enumerate(ps):
        assert p.data_ptr() == opt.state[p]["exp_avg"].data_ptr() - opt._arenas[0].exp_avg.data_ptr() + opt._arenas[0].params.data_ptr()
        assert close(p.data.cpu(), want_p[i]) and update_close(p.data.cpu(), want_p[i], params[i]), i
        assert close(opt.state[p]["exp_avg"].cpu(), want_m[i]) and close(opt.state[p]["exp_avg_sq"].cpu(), want_v[i]), i
        assert float(opt.state[p]["step"]) == len(grads)
    if max_norm and scale == 1e-4:
        assert all(n < max_norm for n in norms)  # below the threshold the coefficient clamps to 1: no scaling


def test_state_dict_moves_to_and_from_torch():
    from dyffusion_b200.optim import AdamW
    params, grads = case(steps=4)
    *_, sd_torch = OO.reference_steps(params, grads[:2], max_grad_norm=1.0, **HYPER)
    want_p, want_m, want_v, _, _ = OO.reference_steps(params, grads, max_grad_norm=1.0, **HYPER)
    mid_p, *_ = OO.reference_steps(params, grads[:2], max_grad_norm=1.0, **HYPER)
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in mid_p]
    opt = AdamW(ps, max_grad_norm=1.0, **HYPER)
    sd_torch["param_groups"][0]["max_grad_norm"] = 1.0
    opt.load_state_dict(sd_torch)           # resume a torch.optim.AdamW run after two steps
    for gs in grads[2:]:
        opt.zero_grad()
        for p, g in zip(ps, gs):
            p.grad.copy_(g.cuda())
        opt.step()
    for i, p in enumerate(ps):
        assert close(p.data.cpu(), want_p[i]) and close(opt.state[p]["exp_avg_sq"].cpu(), want_v[i]), i
        assert update_close(p.data.cpu(), want_p[i], params[i]), i
    sd = opt.state_dict()                   # ... and hand the state back to torch
    back = torch.optim.AdamW([torch.nn.Parameter(p.detach().cpu().clone()) for p in ps], **HYPER)
    sd["param_groups"][0].pop("max_grad_norm")
    back.load_state_dict(sd)
    st = back.state[back.param_groups[0]["params"][0]]
    assert float(st["step"]) == 4 and torch.allclose(st["exp_avg"], want_m[0], rtol=1e-5, atol=1e-9)


def test_engine_backbone_follows_the_arena():
    """The backbone's nn.Parameters become arena views; after a step the engine re-packs its weights (`owners`)."""
    from dyffusion_b200.optim import AdamW
    from tests import helpers as H
    from tests.gpu_helpers import build_backbone
    net = build_backbone("spring", "F", seed=1)
    x, cond = H.forward_inputs("spring", "F", rows=2)
    t = torch.tensor([1.0, 2.0]).cuda()
    with torch.no_grad():
        y0 = net(x.cuda(), time=t, condition=cond.cuda())
    opt = AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-2, betas=(0.9, 0.99), weight_decay=0.0, owners=[net])
    with torch.no_grad():
        y1 = net(x.cuda(), time=t, condition=cond.cuda())
    assert torch.equal(y0, y1)              # moving the parameters into the arena changes nothing
    for p in net.parameters():
        p.grad.fill_(1.0)
    opt.step()
    with torch.no_grad():
        y2 = net(x.cuda(), time=t, condition=cond.cuda())
    assert not torch.equal(y1, y2) and torch.isfinite(y2).all()
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    with torch.no_grad():
        y_ref = H.oracle_net("spring", "F", sd)(x, t.cpu(), cond)
    assert H.rel_l2(y2.cpu(), y_ref) <= 1e-2  # the forward uses the UPDATED weights
