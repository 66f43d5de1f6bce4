from .unet import UNet, Unet
