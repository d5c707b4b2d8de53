from .simple_conv_net import SimpleConvNet  # noqa: F401
from .unet import Unet  # noqa: F401
from .unet_simple import UNet  # noqa: F401
