"""The algebra the fused decoder block rests on (csrc/conv_up.cu) and the backward formulation planned on top of it (DESIGN.md
section 10), checked in fp64 against torch itself: upsample -> conv forward, and its input / weight gradients by autograd."""
import pytest
import torch
import torch.nn.functional as F

from oracle import composite_math as M


def _case(H, W, R=2, Cin=3, Cout=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, Cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(R, Cout, 2 * H, 2 * W, generator=g, dtype=torch.float64)
    y = F.conv2d(F.interpolate(x, scale_factor=2, mode="bilinear"), w, padding=1)  # unet_simple.py:41-51, :101
    y.backward(dy)
    return x, w, dy, y


@pytest.mark.parametrize("H,W", [(4, 5), (2, 2), (6, 3)])
def test_composite_forward_and_backward_equal_torch(H, W):
    x, w, dy, y = _case(H, W)
    with torch.no_grad():
        assert torch.allclose(M.forward_composite(x, w), y, rtol=1e-12, atol=1e-12)
        assert torch.allclose(M.dgrad_composite(dy, w, H, W), x.grad, rtol=1e-12, atol=1e-12)
        assert torch.allclose(M.wgrad_composite(x, dy), w.grad, rtol=1e-12, atol=1e-12)


def test_axis_maps_interior_is_the_fixed_stencil():
    m = M.axis_maps(1)
    # output 2i (a = 0): taps read u[2i-1], u[2i], u[2i+1] = (3/4 x[i-1] + 1/4 x[i]), (1/4 x[i-1] + 3/4 x[i]), (3/4 x[i] + 1/4 x[i+1])
    assert m[0].tolist() == [[0.75, 0.25, 0.0], [0.25, 0.75, 0.75], [0.0, 0.0, 0.25]]
    assert torch.equal(m[1], m[0].flip(0).flip(1))  # output 2i+1 is the mirror image
    assert float(M.axis_maps(0)[0, 0].abs().sum()) == 0.0 and float(M.axis_maps(2)[1, 2].abs().sum()) == 0.0  # nothing beyond the edge
