"""fp64 checks of the backward formulas planned for the CUDA backward kernels (oracle/backward_math.py)
against torch.autograd through the reference's layer pattern conv -> BatchNorm(batch statistics) -> time scale/shift -> act."""
import pytest
import torch
import torch.nn.functional as F

from oracle import backward_math as B


def _rand(*shape, seed=0, grad=False):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64, requires_grad=grad)


def test_layer_pattern_backward_equals_autograd():
    R, Cin, Cout, H, W = 3, 5, 6, 8, 6
    x, w = _rand(R, Cin, H, W, seed=1, grad=True), _rand(Cout, Cin, 3, 3, seed=2, grad=True)
    gamma, beta = _rand(Cout, seed=3, grad=True), _rand(Cout, seed=4, grad=True)
    scale, shift = _rand(R, Cout, 1, 1, seed=5, grad=True), _rand(R, Cout, 1, 1, seed=6, grad=True)
    dy = _rand(R, Cout, H, W, seed=7)
    z = F.conv2d(x, w, padding=1)
    n = F.batch_norm(z, None, None, gamma, beta, training=True, eps=1e-5)
    a = n * (scale + 1) + shift
    y = F.leaky_relu(a, 0.2)
    y.backward(dy)
    with torch.no_grad():
        da = dy * torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.2))
        dn, dscale, dshift = B.scale_shift_backward(n, da, scale)
        dz, dgamma, dbeta = B.bn_train_backward(z, dn, gamma)
        close = lambda u, v: torch.allclose(u, v, rtol=1e-10, atol=1e-12)
        assert close(dscale, scale.grad) and close(dshift, shift.grad)
        assert close(dgamma, gamma.grad) and close(dbeta, beta.grad)
        assert close(B.conv3x3_dgrad_as_conv(dz, w), x.grad)
        assert close(B.conv_wgrad_as_gemm(x, dz, 3, 1, 1), w.grad)


@pytest.mark.parametrize("H,W", [(8, 6), (4, 4), (2, 10)])
def test_stride2_dgrad_by_parity_and_wgrad_equal_autograd(H, W):
    R, Cin, Cout = 2, 3, 4
    x, w = _rand(R, Cin, H, W, seed=8, grad=True), _rand(Cout, Cin, 4, 4, seed=9, grad=True)
    z = F.conv2d(x, w, stride=2, padding=1)  # unet_simple.py:30-33 encoder conv
    dz = _rand(*z.shape, seed=10)
    z.backward(dz)
    with torch.no_grad():
        assert torch.allclose(B.conv4x4s2_dgrad_by_parity(dz, w, H, W), x.grad, rtol=1e-10, atol=1e-12)
        assert torch.allclose(B.conv_wgrad_as_gemm(x, dz, 4, 2, 1), w.grad, rtol=1e-10, atol=1e-12)
