"""DynUNet fixture cases shared by make_golden.py (reference side) and the tests: (constructor kwargs, input shape, weight seed)."""
DYNUNET_CASES = [
    # isotropic nnU-Net default (affine instance norm)
    (dict(spatial_dims=3, in_channels=1, out_channels=2, kernel_size=[3, 3, 3, 3], strides=[1, 2, 2, 2], upsample_kernel_size=[2, 2, 2]), (2, 1, 16, 32, 32), 20),
    # anisotropic kernels / strides, residual blocks, explicit filters
    (dict(spatial_dims=3, in_channels=2, out_channels=3, kernel_size=[[3, 3, 3], [3, 3, 3], [1, 3, 3], [3, 3, 3]],
          strides=[[1, 1, 1], [2, 2, 2], [1, 2, 2], [2, 2, 2]], upsample_kernel_size=[[2, 2, 2], [1, 2, 2], [2, 2, 2]], res_block=True,
          filters=[8, 16, 24, 32]), (1, 2, 16, 32, 24), 21),
    # deep supervision (eval mode: the heads are parameters only), plain instance norm, transposed-conv bias
    (dict(spatial_dims=3, in_channels=1, out_channels=2, kernel_size=[3, 3, 3, 3], strides=[1, 2, 2, 2], upsample_kernel_size=[2, 2, 2],
          deep_supervision=True, deep_supr_num=2, norm_name="instance", trans_bias=True, filters=[16, 32, 48, 64]), (1, 1, 32, 32, 32), 22),
]
