"""GPU: monai_b200.networks.layers.AffineTransform (one host matrix + b200_resample_affine per batch item) against the reference's
unit-test goldens (tests/networks/layers/test_affine_transform.py) and against torch's own F.affine_grid + F.grid_sample on the GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from monai_b200.networks.layers import AffineTransform
from oracle import transforms as otr
from test_host_transforms import AFFINE_TRANSFORM_GOLDENS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", range(len(AFFINE_TRANSFORM_GOLDENS)))
def test_affine_transform_reference_unit_goldens_on_the_gpu(case):
    init, image, theta, call_size, expected, atol = AFFINE_TRANSFORM_GOLDENS[case]
    image = torch.as_tensor(np.asarray(image), dtype=torch.float32).cuda()
    theta = torch.as_tensor(np.asarray(theta), dtype=torch.float32).cuda()
    out = AffineTransform(**init)(image, theta, call_size)
    assert out.dtype == image.dtype and out.is_cuda
    np.testing.assert_allclose(out.cpu().numpy(), np.asarray(expected), atol=max(atol, 1e-4), rtol=1e-4)


@pytest.mark.parametrize("normalized,reverse,align,zero", [(False, True, True, False), (False, True, False, True), (False, False, False, False),
                                                           (True, False, False, False), (True, True, True, False)])
@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
def test_affine_transform_matches_the_restated_reference_chain(normalized, reverse, align, zero, padding):
    """Random batched theta (one per item), 3-D, output size != input size, against oracle.transforms.affine_transform (the reference's
    forward restated on torch CPU ops); fp32 tolerance 1e-4 of the value range."""
    gen = torch.Generator().manual_seed(7)
    src, dst = (9, 12, 10), (11, 8, 13)
    th = torch.eye(4).repeat(2, 1, 1)
    th[:, :3, :3] += torch.randn((2, 3, 3), generator=gen) * 0.2
    th[:, :3, 3] = torch.randn((2, 3), generator=gen) * (0.2 if normalized else 1.5)
    img = torch.randn((2, 3, *src), generator=gen)
    want = otr.affine_transform(img, th, spatial_size=dst, normalized=normalized, mode="bilinear", padding_mode=padding, align_corners=align,
                                reverse_indexing=reverse, zero_centered=zero).numpy()
    layer = AffineTransform(dst, normalized=normalized, mode="bilinear", padding_mode=padding, align_corners=align, reverse_indexing=reverse,
                            zero_centered=None if normalized else zero)
    got = layer(img.cuda(), th.cuda()).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-4)
    got34 = layer(img.cuda(), th[:, :3].cuda()).cpu().numpy()      # N x 3 x 4 thetas are padded to homogeneous form
    np.testing.assert_array_equal(got34, got)


def test_affine_transform_forward_2d_identity_of_the_reference_test():
    """test_forward_2d (test_affine_transform.py:332-352): normalized=True, reverse_indexing=False is plain affine_grid + grid_sample."""
    x = torch.rand(2, 1, 4, 4, generator=torch.Generator().manual_seed(1))
    theta = torch.tensor([[[0.0, -1.0, 0.0], [1.0, 0.0, 0.0]]]).repeat(2, 1, 1)
    expected = F.grid_sample(x, F.affine_grid(theta, x.size(), align_corners=False), align_corners=False).numpy()
    for th in (theta, theta[0], theta[:1]):
        got = AffineTransform(normalized=True, reverse_indexing=False, align_corners=False)(x.cuda(), th.cuda()).cpu().numpy()
        np.testing.assert_allclose(got, expected, rtol=1e-5, atol=1e-5)
