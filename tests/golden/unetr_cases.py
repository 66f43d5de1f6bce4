"""UNETR fixture cases shared by make_golden.py (reference side) and the tests: (constructor kwargs, oracle kwargs, input shape, seed)."""
UNETR_CASES = [
    # small ViT (4 heads of 24), instance norm, residual blocks: the reference's structure at a size the CPU oracle runs in seconds
    (dict(in_channels=1, out_channels=2, img_size=(32, 32, 32), feature_size=16, hidden_size=96, mlp_dim=192, num_heads=4), dict(num_heads=4), (2, 1, 32, 32, 32), 60),
    # anisotropic image, batch norm (eval), plain conv blocks, qkv bias, 8 heads of 8
    (dict(in_channels=2, out_channels=3, img_size=(32, 48, 32), feature_size=16, hidden_size=64, mlp_dim=128, num_heads=8, norm_name="batch",
          res_block=False, qkv_bias=True), dict(num_heads=8, res_block=False), (1, 2, 32, 48, 32), 61),
]
