"""dyffusion_b200 -- B200-native DYffusion sampling engine behind the reference's nn.Module / Hydra surface.

Drop-in `_target_`s:
    model:     dyffusion_b200.backbones.unet_simple.UNet | dyffusion_b200.backbones.unet.Unet |
               dyffusion_b200.backbones.simple_conv_net.SimpleConvNet
    diffusion: dyffusion_b200.diffusion.dyffusion.DYffusion
The arithmetic runs in hand-written sm_100a CUDA behind the C ABI of include/dyffusion_b200.h.
"""
__version__ = "0.1.0"
