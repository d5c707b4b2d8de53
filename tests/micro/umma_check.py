import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from dyffusion_b200.backbones import SimpleConvNet
from oracle.synth import synth_state_dict, synth_tensor
from oracle import dyffusion_oracle as O
from tests import helpers as H

def run(hw, ks, rows=2, dim=64):
    m = SimpleConvNet(dim=dim, with_time_emb=True, kernel_sizes=ks, dropout=0.0, num_input_channels=4,
                      num_output_channels=4, num_conditional_channels=1, spatial_shape=hw, verbose=False)
    sd = synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=5)
    m.load_state_dict(sd); m = m.cuda().eval()
    x = synth_tensor("q.x", (rows, 4, *hw)).cuda(); c = synth_tensor("q.c", (rows, 1, *hw)).cuda()
    t = torch.linspace(0.5, 2.0, rows).cuda()
    with torch.no_grad():
        os.environ["DYF_DISABLE_UMMA"] = "1"
        y_mma = m(x, time=t, condition=c)
        del os.environ["DYF_DISABLE_UMMA"]
        y_umma = m(x, time=t, condition=c)
        y_or = O.simple_conv_net_forward(sd, x.cpu(), t.cpu(), c.cpu(), dim=dim, kernel_sizes=ks)
    d = (y_umma - y_mma).abs()
    print(f"hw={hw} ks={ks}: umma~mma {H.rel_l2(y_umma.cpu(), y_mma.cpu()):.2e} | mma~oracle {H.rel_l2(y_mma.cpu(), y_or):.2e} "
          f"| umma~oracle {H.rel_l2(y_umma.cpu(), y_or):.2e} | maxdiff {float(d.max()):.3e} at {tuple(int(i) for i in (d==d.max()).nonzero()[0])}")

for hw in [(16, 8), (32, 16), (10, 10), (40, 24), (60, 60), (17, 9)]:
    run(hw, [3, 3])
run((32, 32), [3, 3, 3, 3])
run((64, 64), [3, 3], dim=128)

def run_ns(up, rows=2):
    from dyffusion_b200.backbones import UNet
    m = UNet(dim=64, with_time_emb=True, upsample_dims=up, dropout=0.0, num_input_channels=3, num_output_channels=3,
             num_conditional_channels=2, spatial_shape=(221, 42), verbose=False)
    sd = synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=6)
    m.load_state_dict(sd); m = m.cuda().eval()
    x = synth_tensor("n.x", (rows, 3, 221, 42)).cuda(); c = synth_tensor("n.c", (rows, 2, 221, 42), kind="mask").cuda()
    t = torch.linspace(0.0, 7.0, rows).cuda()
    with torch.no_grad():
        os.environ["DYF_DISABLE_UMMA"] = "1"
        y_mma = m(x, time=t, condition=c)
        del os.environ["DYF_DISABLE_UMMA"]
        y_umma = m(x, time=t, condition=c)
        y_or = O.unet_simple_forward(sd, x.cpu(), t.cpu(), c.cpu(), dim=64, upsample_dims=up)
    print(f"NS up={up}: umma~mma {H.rel_l2(y_umma.cpu(), y_mma.cpu()):.2e} | mma~oracle {H.rel_l2(y_mma.cpu(), y_or):.2e} "
          f"| umma~oracle {H.rel_l2(y_umma.cpu(), y_or):.2e}")

run_ns((64, 64)); run_ns((128, 64)); run_ns((256, 256))
