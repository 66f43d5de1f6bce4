from .convutils import gaussian_1d, same_padding, stride_minus_kernel_padding
from .factories import get_act_layer, get_dropout_layer, get_norm_layer, split_args
from .spatial_transforms import AffineTransform, grid_count, grid_grad, grid_pull, grid_push
