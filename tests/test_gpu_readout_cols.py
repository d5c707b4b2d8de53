"""NS readout on the sampled source columns only (net.cu: `ro_xmap`): the ConvTranspose2d channel contraction runs as a 1x1
tcgen05 conv over the 126 of 256 source columns that the outer 512 -> 42 bilinear resize reads, gathered through a column
table (reference: src/models/unet_simple.py:141-150 readout, :195 outer resize).  The full-width conv over a TMA box
(`DYF_DISABLE_READOUT_COLS=1` at net creation) is the in-engine reference: both accumulate the same 64 products per output
in the same order, so the two networks must agree bit for bit; the oracle comparison is tests/test_gpu_parity.py."""
import os

import pytest
import torch

from oracle.synth import synth_state_dict, synth_tensor

pytestmark = pytest.mark.gpu


def _build(cols, spatial, up):
    from dyffusion_b200.backbones import UNet
    saved = os.environ.pop("DYF_DISABLE_READOUT_COLS", None)
    try:
        if not cols:
            os.environ["DYF_DISABLE_READOUT_COLS"] = "1"
        m = UNet(dim=64, with_time_emb=True, upsample_dims=up, dropout=0.15, num_input_channels=3, num_output_channels=3,
                 num_conditional_channels=2, spatial_shape=spatial, verbose=False)
    finally:
        os.environ.pop("DYF_DISABLE_READOUT_COLS", None)
        if saved is not None:
            os.environ["DYF_DISABLE_READOUT_COLS"] = saved
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=9))
    return m.cuda().eval()


@pytest.mark.parametrize("spatial,up,rows", [((221, 42), [256, 256], 3), ((221, 42), [128, 128], 2), ((37, 19), [64, 64], 5)])
def test_column_subset_readout_is_bit_exact(spatial, up, rows):
    a, b = _build(True, spatial, up), _build(False, spatial, up)
    x = synth_tensor("ro.x", (rows, 3, *spatial)).cuda()
    c = synth_tensor("ro.c", (rows, 2, *spatial), kind="mask").cuda()
    t = torch.arange(rows, dtype=torch.float32).cuda()
    with torch.no_grad():
        ya, yb = a(x, time=t, condition=c), b(x, time=t, condition=c)
        with a.inference_dropout_scope(True), b.inference_dropout_scope(True):
            da, db = a(x, time=t, condition=c), b(x, time=t, condition=c)
    assert torch.isfinite(ya).all() and ya.shape == (rows, 3, *spatial)
    assert torch.equal(ya, yb)
    assert torch.equal(da, db)
