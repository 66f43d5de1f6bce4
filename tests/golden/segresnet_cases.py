"""SegResNet fixture cases shared by make_golden.py (reference side) and the tests:
(constructor kwargs, oracle kwargs, input shape, weight seed)."""
SEGRESNET_CASES = [
    # the reference's defaults: GroupNorm(8), ReLU, trilinear upsampling, blocks (1,2,2,4) / (1,1,1)
    (dict(), dict(), (2, 1, 16, 32, 24), 30),
    # transposed-convolution upsampling, 4 groups, multi-channel input
    (dict(spatial_dims=3, init_filters=16, in_channels=4, out_channels=3, blocks_down=(1, 2, 2), blocks_up=(1, 1), upsample_mode="deconv",
          norm=("GROUP", {"num_groups": 4})), dict(blocks_down=(1, 2, 2), blocks_up=(1, 1), groups=4, upsample_mode="deconv"), (1, 4, 16, 16, 24), 31),
    # instance norm + LeakyReLU (bundles that override norm / act), dropout inactive in eval mode
    (dict(init_filters=8, norm="instance", dropout_prob=0.2, act=("leakyrelu", {"negative_slope": 0.1})), dict(groups=0, slope=0.1), (1, 1, 16, 16, 16), 32),
]
