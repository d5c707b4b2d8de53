"""TEST INFRASTRUCTURE ONLY -- fp64 statements of the backward formulas the CUDA backward kernels will implement (DESIGN.md
section 10), each checked against torch.autograd in tests/test_backward_math_cpu.py.  Only tests/ may import this.

Layer pattern of the Navier-Stokes / spring-mesh backbones in train mode (reference src/models/unet_simple.py:13-83,
src/models/simple_conv_net.py:12-56):  z = conv(x);  n = BatchNorm_batch(z);  a = n * (scale + 1) + shift;  y = act(a) * mask."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F


def bn_train_backward(z: torch.Tensor, dn: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """n = gamma * (z - mean) / sqrt(var + eps) + beta with batch statistics over (rows, H, W).
    -> (dz, dgamma, dbeta).  Needs exactly two per-channel sums of the incoming gradient: sum(dn) and sum(dn * zhat) -- what
    the wgrad / dgrad prologue accumulates in a fixed order."""
    dims = (0, 2, 3)
    m = z.shape[0] * z.shape[2] * z.shape[3]
    mean = z.mean(dims, keepdim=True)
    var = z.var(dims, unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    zhat = (z - mean) * rstd
    s1 = dn.sum(dims, keepdim=True)              # = dbeta
    s2 = (dn * zhat).sum(dims, keepdim=True)     # = dgamma
    g = gamma.view(1, -1, 1, 1)
    dz = g * rstd * (dn - s1 / m - zhat * s2 / m)
    return dz, s2.flatten(), s1.flatten()


def scale_shift_backward(n: torch.Tensor, da: torch.Tensor, scale: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """a = n * (scale + 1) + shift, scale / shift [rows, C, 1, 1] from the time MLP.  -> (dn, dscale, dshift): the per-(row,
    channel) sums that feed the time-MLP backward."""
    return da * (scale + 1), (da * n).sum((2, 3), keepdim=True), da.sum((2, 3), keepdim=True)


def conv4x4s2_dgrad_by_parity(dz: torch.Tensor, w: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Input gradient of the encoder's conv(k=4, stride=2, pad=1) (unet_simple.py:120-127) WITHOUT a scatter: input pixel
    (2i+a, 2j+b) receives dz[i + (a + 1 - ky) / 2, ...] only from the two ky with ky = a + 1 (mod 2), so each of the four input
    parity classes is a dense 2x2 stride-1 conv over dz -- the mirror image of the parity views the forward reads through.
    dz [R, Cout, H/2, W/2], w [Cout, Cin, 4, 4] -> dx [R, Cin, H, W]."""
    R, Cout, h, wd = dz.shape
    dx = torch.zeros(R, w.shape[1], H, W, dtype=dz.dtype)
    dzp = F.pad(dz, (1, 1, 1, 1))
    for a in range(2):
        for b in range(2):
            # taps with ky = a + 1 (mod 2): ky in {1, 3} for a = 0 -> output rows i, i - 1; ky in {0, 2} for a = 1 -> rows i + 1, i
            kys = [1, 3] if a == 0 else [0, 2]
            kxs = [1, 3] if b == 0 else [0, 2]
            acc = torch.zeros(R, w.shape[1], h, wd, dtype=dz.dtype)
            for ky in kys:
                oy = (a + 1 - ky) // 2  # output row offset relative to i: 2*(i+oy) + ky - 1 = 2i + a
                for kx in kxs:
                    ox = (b + 1 - kx) // 2
                    g = dzp[:, :, 1 + oy:1 + oy + h, 1 + ox:1 + ox + wd]
                    acc += torch.einsum("rohw,oc->rchw", g, w[:, :, ky, kx])
            dx[:, :, a::2, b::2] = acc
    return dx


def conv3x3_dgrad_as_conv(dz: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Input gradient of conv(k=3, stride=1, pad=1) = the same conv with the filter rotated by 180 degrees and Cin <-> Cout
    swapped: the forward tcgen05 kernel with a second weight pack."""
    return F.conv2d(dz, w.flip(2, 3).transpose(0, 1), padding=1)


def conv_wgrad_as_gemm(x: torch.Tensor, dz: torch.Tensor, k: int, stride: int, pad: int) -> torch.Tensor:
    """Weight gradient as ONE GEMM: M = Cout, N = Cin * k * k, K = rows * out pixels; A = dz read pixel-major, B = the im2col
    view of x read pixel-major (both operands MN-major for this contraction)."""
    cols = F.unfold(x, k, padding=pad, stride=stride)            # [R, Cin*k*k, P]
    g = dz.flatten(2)                                            # [R, Cout, P]
    return torch.einsum("rop,rnp->on", g, cols).view(dz.shape[1], x.shape[1], k, k)
