"""Developer tool (GPU): fused upsample+conv decoder blocks (conv_up.cu) against the two-kernel path and the oracle."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from oracle import dyffusion_oracle as O  # noqa: E402
from oracle.synth import synth_state_dict, synth_tensor  # noqa: E402
from tests import helpers as H  # noqa: E402


def build(up, fuse, fuse_min=None):
    from dyffusion_b200.backbones import UNet
    os.environ.pop("DYF_DISABLE_UPFUSE", None)
    os.environ.pop("DYF_UPFUSE_MIN", None)
    if not fuse:
        os.environ["DYF_DISABLE_UPFUSE"] = "1"
    if fuse_min:
        os.environ["DYF_UPFUSE_MIN"] = str(fuse_min)
    m = UNet(dim=64, with_time_emb=True, upsample_dims=up, dropout=0.0, num_input_channels=3, num_output_channels=3,
             num_conditional_channels=2, spatial_shape=(221, 42), verbose=False)
    sd = synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=6)
    m.load_state_dict(sd)
    return m.cuda().eval(), sd


def run(up, rows=3, fuse_min=None):
    x = synth_tensor("n.x", (rows, 3, 221, 42)).cuda()
    c = synth_tensor("n.c", (rows, 2, 221, 42), kind="mask").cuda()
    t = torch.linspace(0.0, 7.0, rows).cuda()
    with torch.no_grad():
        mf, sd = build(up, True, fuse_min)
        yf = mf(x, time=t, condition=c)
        mu, _ = build(up, False)
        yu = mu(x, time=t, condition=c)
        yo = O.unet_simple_forward(sd, x.cpu(), t.cpu(), c.cpu(), dim=64, upsample_dims=up)
    d = (yf - yu).abs()
    print(f"up={up} min={fuse_min}: fused~unfused {H.rel_l2(yf.cpu(), yu.cpu()):.2e} | fused~oracle {H.rel_l2(yf.cpu(), yo):.2e} | "
          f"unfused~oracle {H.rel_l2(yu.cpu(), yo):.2e} | maxdiff {float(d.max()):.3e}", flush=True)


if __name__ == "__main__":
    run((256, 256))
    run((128, 192), fuse_min=32)
    run((64, 64), fuse_min=32)
    run((256, 256), fuse_min=32)
