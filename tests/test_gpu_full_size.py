"""Full-size checks at BASELINE.json's configurations (C2: 256^3, C3: 512^3, C4: 1x256^3 fp32) through size-independent
properties -- the CPU oracle cannot run these sizes in seconds, so parity is anchored on identities the algorithm must
satisfy exactly or to rounding: partition of unity, gather/blend identity, one-shot == streaming, transform idempotence."""
import numpy as np
import pytest
import torch

import monai_b200.inferers.utils as U
from monai_b200.data import MetaTensor
from monai_b200.inferers import SlidingWindowInferer
from monai_b200.networks.layers.convutils import gaussian_1d
from monai_b200.transforms import GaussianSmooth, RandAffine, Spacing

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("vol,nwin", [((256, 256, 256), 125), ((512, 512, 512), 1000)])
def test_partition_of_unity_and_identity_at_full_size(vol, nwin):
    """Every window predicts (i) a constant per channel and (ii) its own input: the Gaussian-weighted blend must return the
    constant / the input volume itself, whatever the 1..27-fold coverage of a voxel (C2 / C3 window tables, fp16)."""
    calls = []

    def pred(w):
        calls.append(w.shape[0])
        const = torch.full_like(w, 1.5)
        return torch.cat([const, w, -0.25 * torch.ones_like(w)], dim=1)

    x = torch.randn((1, 1, *vol), device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)).half()
    y = SlidingWindowInferer((96, 96, 96), 25, 0.5, "gaussian")(x, pred)
    assert sum(calls) == nwin and tuple(y.shape) == (1, 3, *vol) and y.dtype == torch.float16
    assert torch.all(y[:, 0] == 1.5) and torch.all(y[:, 2] == -0.25)          # sum(w c) / sum(w) rounds back to c in fp16
    err = (y[:, 1].float() - x[:, 0].float()).abs()
    tol = x[:, 0].float().abs() * 2.0 ** -10 + 1e-7                            # at most one fp16 ulp
    assert bool((err <= tol).all()), float(err.max())


def test_streaming_equals_one_shot_at_c2_size(monkeypatch):
    """C2 volume with a resident budget of 40 windows: accumulate (mode 1) + finalise (mode 2) must reproduce the one-shot
    blend bit for bit (same fused multiply-add order)."""
    def pred(w):
        return torch.cat([torch.tanh(w), w * w - 0.5], dim=1)

    x = torch.randn((1, 1, 256, 256, 256), device=DEV, generator=torch.Generator(device=DEV).manual_seed(2)).half()
    inf = SlidingWindowInferer((96, 96, 96), 25, 0.5, "gaussian")
    a = inf(x, pred)
    monkeypatch.setattr(U, "_RESIDENT_BYTES", 40 * 2 * 96**3 * 2)
    b = inf(x, pred)
    assert torch.equal(a, b)


def test_transform_identities_at_c4_size():
    """1 x 256^3 fp32: Spacing to the volume's own pixdim and RandAffine with empty ranges are the identity map (bilinear
    weights collapse to 1 at integer coordinates); GaussianSmooth of a constant volume is the constant times the product of
    the truncated tap sums in the interior."""
    g = torch.Generator(device=DEV).manual_seed(3)
    img = torch.rand((1, 256, 256, 256), device=DEV, generator=g)
    m = MetaTensor(img, affine=torch.diag(torch.tensor([1.25, 1.25, 1.25, 1.0], dtype=torch.float64)))
    same = Spacing(pixdim=(1.25, 1.25, 1.25), mode="bilinear")(m)
    assert tuple(same.shape) == (1, 256, 256, 256)
    torch.testing.assert_close(same.as_subclass(torch.Tensor), img, rtol=0, atol=1e-6)
    t = RandAffine(prob=1.0, mode="bilinear", padding_mode="border")
    t.set_random_state(seed=0)
    torch.testing.assert_close(t(m).as_subclass(torch.Tensor), img, rtol=0, atol=1e-6)
    c = torch.full((1, 256, 256, 256), 2.0, device=DEV)
    s = GaussianSmooth(sigma=1.0)(c)
    taps = gaussian_1d(torch.tensor(1.0), truncated=4.0, approx="erf")
    k = float(taps.sum())
    np.testing.assert_allclose(s[0, 8:-8, 8:-8, 8:-8].cpu().numpy(), 2.0 * k**3, rtol=1e-5)
    half = float(taps[taps.numel() // 2:].sum())       # zero padding: the corner voxel sees the centre tap and one tail per axis
    np.testing.assert_allclose(float(s[0, 0, 0, 0]), 2.0 * half**3, rtol=1e-5)
