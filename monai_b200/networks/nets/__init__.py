from .basic_unet import BasicUNet, BasicUnet, Basicunet, basicunet
from .unet import UNet, Unet
