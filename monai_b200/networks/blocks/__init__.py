from .acti_norm import ADN
from .convolutions import Convolution, ResidualUnit
