"""TEST INFRASTRUCTURE ONLY -- fp64 statement of the algebra behind the fused decoder block (`csrc/conv_up.cu`) and of the
backward formulation planned on top of it.  Only tests/ may import this.

Decoder block of `unet_simple` (reference src/models/unet_simple.py:41-51, :101): y = conv3x3_pad1(bilinear_x2(x)).  Bilinear x2
with align_corners=False is the fixed stencil u[2i] = 1/4 x[i-1] + 3/4 x[i], u[2i+1] = 3/4 x[i] + 1/4 x[i+1] (indices clamped),
so for an output pixel (2i+a, 2j+b) the block is a 3x3 conv over the LOW-resolution x with composite weights
    Wc[a,b][di,dj] = sum_{ky,kx} V_a[di,ky] H_b[dj,kx] w[ky,kx],
where the 1-D maps V / H depend on whether i (j) is the first, an interior or the last index of its axis (the reference clamps
the interpolation but zero-pads the upsampled map).  `axis_maps` restates the C++ function of the same name."""
from __future__ import annotations

import torch


def axis_maps(where: int) -> torch.Tensor:
    """M[a, d, k] = coefficient of w[k] * x[i + d - 1] in output 2i + a, for i first (0) / interior (1) / last (2)."""
    L = 5
    i = 0 if where == 0 else 2 if where == 1 else L - 1
    M = torch.zeros(2, 3, 3, dtype=torch.float64)
    for a in range(2):
        for d in range(3):
            for k in range(3):
                src, r = i + d - 1, 2 * i + a + k - 1
                if 0 <= src < L and 0 <= r < 2 * L:
                    q = r >> 1
                    lo = q if r & 1 else max(q - 1, 0)
                    hi = min(q + 1, L - 1) if r & 1 else q
                    wlo = 0.75 if r & 1 else 0.25
                    M[a, d, k] = (wlo if lo == src else 0.0) + ((1 - wlo) if hi == src else 0.0)
    return M


def _where(i: int, n: int) -> int:
    return 0 if i == 0 else 2 if i == n - 1 else 1


def composite_weights(w: torch.Tensor, wy: int, wx: int) -> torch.Tensor:
    """w [Cout, Cin, 3, 3] -> Wc [2, 2, Cout, Cin, 3, 3] for the (row, column) position classes (wy, wx)."""
    V, Hm = axis_maps(wy), axis_maps(wx)
    return torch.einsum("aik,bjl,ockl->abocij", V, Hm, w.double())


def forward_composite(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x [R, Cin, H, W] -> [R, Cout, 2H, 2W]: the block evaluated on the low-resolution grid, pixel by pixel."""
    R, Cin, H, W = x.shape
    xp = torch.nn.functional.pad(x.double(), (1, 1, 1, 1))
    y = torch.zeros(R, w.shape[0], 2 * H, 2 * W, dtype=torch.float64)
    for i in range(H):
        for j in range(W):
            Wc = composite_weights(w, _where(i, H), _where(j, W))
            patch = xp[:, :, i:i + 3, j:j + 3]
            y[:, :, 2 * i:2 * i + 2, 2 * j:2 * j + 2] = torch.einsum("rcij,abocij->roab", patch, Wc)
    return y


def dgrad_composite(dy: torch.Tensor, w: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """dy [R, Cout, 2H, 2W] -> dx [R, Cin, H, W]: the transposed composite conv, again entirely on the low-resolution grid --
    the upsampled tensor (and its gradient) never exists.  Interior pixels share one weight set: there the sum below is a
    plain 3x3 conv over space_to_depth(dy) (4*Cout channels) with Wc transposed and flipped."""
    R = dy.shape[0]
    dxp = torch.zeros(R, w.shape[1], H + 2, W + 2, dtype=torch.float64)
    for i in range(H):
        for j in range(W):
            Wc = composite_weights(w, _where(i, H), _where(j, W))
            g = dy[:, :, 2 * i:2 * i + 2, 2 * j:2 * j + 2].double()
            dxp[:, :, i:i + 3, j:j + 3] += torch.einsum("roab,abocij->rcij", g, Wc)
    return dxp[:, :, 1:-1, 1:-1]


def wgrad_composite(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """-> dw [Cout, Cin, 3, 3]: the composite-weight gradients dWc (one GEMM over pixels per position class, the shape the
    tcgen05 wgrad kernel produces) mapped back to the 3x3 filter with the transposed axis maps (a 9 x 36 linear map)."""
    R, Cin, H, W = x.shape
    xp = torch.nn.functional.pad(x.double(), (1, 1, 1, 1))
    dw = torch.zeros(dy.shape[1], Cin, 3, 3, dtype=torch.float64)
    for wy in range(3):
        for wx in range(3):
            dWc = torch.zeros(2, 2, dy.shape[1], Cin, 3, 3, dtype=torch.float64)
            for i in [k for k in range(H) if _where(k, H) == wy]:
                for j in [k for k in range(W) if _where(k, W) == wx]:
                    g = dy[:, :, 2 * i:2 * i + 2, 2 * j:2 * j + 2].double()
                    dWc += torch.einsum("roab,rcij->abocij", g, xp[:, :, i:i + 3, j:j + 3])
            dw += torch.einsum("aik,bjl,abocij->ockl", axis_maps(wy), axis_maps(wx), dWc)
    return dw
