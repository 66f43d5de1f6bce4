from .acti_norm import ADN
from .convolutions import Convolution, ResidualUnit
from .dynunet_block import UnetBasicBlock, UnetOutBlock, UnetResBlock, UnetUpBlock, get_conv_layer, get_output_padding, get_padding
from .segresnet_block import ResBlock, UpSample
from .transformerblock import MLPBlock, PatchEmbeddingBlock, SABlock, TransformerBlock
from .unetr_block import UnetrBasicBlock, UnetrPrUpBlock, UnetrUpBlock
