"""GPU parity of the transform operator surface (Spacing / Spacingd / RandAffined / GaussianSmooth) vs the reference
fixtures and the torch-CPU oracle.  fp32 resample tolerance 1e-3 relative (north star); measured errors are ~1e-5."""
import ast
import itertools
import os

import numpy as np
import pytest
import torch

from monai_b200.data import MetaTensor
from monai_b200.transforms import (Activations, Activationsd, AsDiscrete, AsDiscreted, Compose, GaussianSmooth, GaussianSmoothd, RandAffine,
                                   RandAffined, Spacing, Spacingd)
from oracle import transforms as otr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _kw(g, tag):
    return ast.literal_eval(str(g[f"{tag}.kw"]))


def test_spacing_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    for tag in ("s0", "s1", "s2", "s3", "s4"):
        kw = _kw(g, tag)
        m = MetaTensor(torch.from_numpy(g["img"]).to(DEV), affine=torch.as_tensor(g[f"{tag}.affine"]))
        r = Spacing(pixdim=tuple(g[f"{tag}.pixdim"]), **kw)(m)
        assert tuple(r.shape) == g[f"{tag}.y"].shape and r.dtype == torch.float32, tag  # shapes are bit-exact host arithmetic
        np.testing.assert_allclose(r.affine.numpy(), g[f"{tag}.new_affine"], rtol=1e-9, atol=1e-9, err_msg=tag)
        if kw.get("mode") == "nearest":
            assert (r.cpu().numpy() != g[f"{tag}.y"]).mean() < 2e-3, tag
        else:
            np.testing.assert_allclose(r.cpu().numpy(), g[f"{tag}.y"], rtol=1e-3, atol=1e-4, err_msg=tag)


def test_rand_affined_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    for tag in ("r0", "r1", "r2"):
        kw = _kw(g, tag)
        t = RandAffined(keys=["image"], **kw)
        t.set_random_state(seed=0)
        r = t({"image": MetaTensor(torch.from_numpy(g["img2"]).to(DEV), affine=torch.eye(4))})["image"]
        assert tuple(r.shape) == g[f"{tag}.y"].shape, tag
        if kw["mode"] == "nearest":
            assert (r.cpu().numpy() != g[f"{tag}.y"]).mean() < 2e-3, tag
        else:
            np.testing.assert_allclose(r.cpu().numpy(), g[f"{tag}.y"], rtol=1e-3, atol=2e-4, err_msg=tag)
        np.testing.assert_allclose(r.affine.numpy(), g[f"{tag}.new_affine"], rtol=1e-5, atol=1e-5, err_msg=tag)


def test_gaussian_smooth_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    img = torch.from_numpy(g["img2"]).to(DEV)
    for tag in ("g0", "g1", "g2"):
        s = g[f"{tag}.sigma"].tolist()
        y = GaussianSmooth(sigma=s[0] if len(s) == 1 else s)(img)
        np.testing.assert_allclose(y.cpu().numpy(), g[f"{tag}.y"], rtol=1e-4, atol=1e-5)
    # the 2-D golden of tests/transforms/test_gaussian_smooth.py
    img2 = torch.tensor([[[1, 1, 1], [2, 2, 2], [3, 3, 3]], [[4, 4, 4], [5, 5, 5], [6, 6, 6]]], dtype=torch.float32, device=DEV)
    exp = [[[0.59167546, 0.69312394, 0.59167546], [0.7956997, 0.93213004, 0.7956997], [0.7668002, 0.8982755, 0.7668002]],
           [[1.6105323, 1.8866735, 1.6105323], [1.9892492, 2.3303251, 1.9892492], [1.7856569, 2.091825, 1.7856569]]]
    np.testing.assert_allclose(GaussianSmooth(sigma=1.5)(img2).cpu().numpy(), np.array(exp), rtol=1e-4, atol=1e-4)


def test_config_c4_pipeline_vs_oracle():
    """Spacingd -> RandAffined -> GaussianSmoothd on a small volume vs the CPU oracle."""
    gen = torch.Generator().manual_seed(5)
    img = torch.rand((1, 40, 36, 44), generator=gen)
    aff = np.diag([1.25, 1.25, 1.25, 1.0])
    pipe = Compose([
        Spacingd(keys=["image"], pixdim=(1.0, 1.0, 1.0), mode="bilinear"),
        RandAffined(keys=["image"], prob=1.0, rotate_range=(0.2,) * 3, scale_range=(0.1,) * 3, translate_range=(5,) * 3, mode="bilinear", padding_mode="border"),
        GaussianSmoothd(keys=["image"], sigma=1.0),
    ])
    pipe.transforms[1].set_random_state(seed=0)
    got = pipe({"image": MetaTensor(img.to(DEV), affine=torch.as_tensor(aff))})["image"]
    a, _ = otr.spacing(img, aff, (1.0, 1.0, 1.0))
    b, _ = otr.rand_affine(a, 0, (0.2,) * 3, (), (5,) * 3, (0.1,) * 3, None, "bilinear", "border")
    want = otr.gaussian_smooth(b, 1.0)
    assert tuple(got.shape) == tuple(want.shape) == (1, 50, 45, 55)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-3, atol=2e-4)


def test_rand_affine_without_transform_and_plain_tensor():
    x = torch.rand((1, 8, 9, 10), device=DEV)
    t = RandAffine(prob=0.0, rotate_range=(0.5,))
    t.set_random_state(seed=1)
    y = t(x)
    torch.testing.assert_close(y, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        Spacing(pixdim=(2.0, 2.0, 2.0))(torch.rand(1, 4, 4, 4))


def test_post_transforms_match_reference_fixture(golden_dir):
    """Activations / AsDiscrete on the channel-wise CUDA kernels vs outputs of the real reference; argmax ties pick the first
    maximum, rounding is half-to-even, results are float32, MetaTensor metadata survives, dict versions and fp16 inputs work."""
    g = np.load(os.path.join(golden_dir, "post.npz"))
    logits = torch.from_numpy(g["logits"]).to(DEV)
    np.testing.assert_allclose(Activations(softmax=True)(logits).cpu().numpy(), g["softmax"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(Activations(sigmoid=True)(logits).cpu().numpy(), g["sigmoid"], rtol=2e-6, atol=1e-7)
    a = AsDiscrete(argmax=True)(logits)
    assert a.dtype == torch.float32 and tuple(a.shape) == (1, 6, 7, 5)
    np.testing.assert_array_equal(a.cpu().numpy(), g["argmax"])
    np.testing.assert_array_equal(AsDiscrete(argmax=True, to_onehot=3)(logits).cpu().numpy(), g["argmax_onehot"])
    np.testing.assert_array_equal(AsDiscrete(threshold=0.25)(logits).cpu().numpy(), g["threshold"])
    r = torch.tensor([[0.5, 1.5, 2.5, -0.5, -1.5, 0.49, 2.51]], device=DEV)
    np.testing.assert_array_equal(AsDiscrete(rounding="torchrounding")(r).cpu().numpy(), g["round"])
    np.testing.assert_array_equal(AsDiscrete(to_onehot=3)(torch.from_numpy(g["labels"]).to(DEV)).cpu().numpy(), g["onehot"])
    d = Compose([Activationsd(keys=["pred"], sigmoid=True), AsDiscreted(keys=["pred"], threshold=0.5)])({"pred": MetaTensor(logits, affine=torch.eye(4) * 2)})
    assert isinstance(d["pred"], MetaTensor) and float(d["pred"].affine[0, 0]) == 2.0
    np.testing.assert_array_equal(d["pred"].cpu().numpy(), g["sigmoid_threshold"])
    # fp16 logits (what the fp16 inferer returns): softmax computed in fp32, returned as fp16; argmax unaffected by the dtype
    h = logits.half()
    np.testing.assert_allclose(Activations(softmax=True)(h).float().cpu().numpy(), otr.activations(h.float().cpu(), softmax=True).numpy(), atol=1e-3)
    np.testing.assert_array_equal(AsDiscrete(argmax=True)(h).cpu().numpy(), otr.as_discrete(h.float().cpu(), argmax=True).numpy())
    # error behaviour of the reference
    with pytest.raises(ValueError, match="Incompatible values"):
        Activations()(logits, sigmoid=True, softmax=True)
    with pytest.raises(TypeError, match="other must be None or callable"):
        Activations(other=3)
    with pytest.raises(ValueError, match="deprecated"):
        AsDiscrete(to_onehot=True)
    np.testing.assert_allclose(Activations(other=torch.tanh)(logits).cpu().numpy(), np.tanh(g["logits"]), rtol=1e-6, atol=1e-6)


def test_lazy_resampling_and_invertd_match_the_real_reference(golden_dir):
    """f1: Compose(lazy=True) -- Spacingd o RandAffined as ONE resample -- and f2: Invertd(Spacingd) of a prediction, values
    against fixtures of the real reference (tests/golden/lazy_inverse.npz)."""
    from monai_b200.data import MetaTensor
    from monai_b200.transforms import Compose, Invertd, RandAffined, Spacingd

    g = np.load(os.path.join(golden_dir, "lazy_inverse.npz"))

    def pipe(lazy):
        c = Compose([Spacingd(keys=["image"], pixdim=(1.0, 1.0, 1.0), mode="bilinear"),
                     RandAffined(keys=["image"], prob=1.0, rotate_range=(0.2,) * 3, scale_range=(0.1,) * 3, translate_range=(5,) * 3, mode="bilinear", padding_mode="border")],
                    lazy=lazy)
        c.transforms[1].set_random_state(seed=0)
        return c

    for tag, lazy in (("eager", False), ("lazy", True)):
        x = MetaTensor(torch.from_numpy(g["x"]).to(DEV), affine=torch.as_tensor(g["x_affine"]))
        y = pipe(lazy)({"image": x})["image"]
        np.testing.assert_allclose(y.cpu().numpy(), g[f"{tag}.y"], rtol=1e-4, atol=1e-4, err_msg=tag)
        np.testing.assert_allclose(np.asarray(y.affine), g[f"{tag}.affine"], atol=1e-6)
    pre = Spacingd(keys=["image"], pixdim=(1.0, 1.0, 1.0), mode="bilinear")
    d = pre({"image": MetaTensor(torch.from_numpy(g["x"]).to(DEV), affine=torch.as_tensor(g["x_affine"]))})
    np.testing.assert_allclose(d["image"].cpu().numpy(), g["pre.y"], rtol=1e-4, atol=1e-4)
    pred = MetaTensor(torch.from_numpy(g["pred"]).to(DEV))
    for tag, nearest in (("nearest", True), ("bilinear", False)):
        inv = Invertd(keys=["pred"], transform=pre, orig_keys=["image"], nearest_interp=nearest)({"image": d["image"], "pred": pred})["pred"]
        assert tuple(inv.shape) == tuple(g[f"inv.{tag}"].shape)
        diff = np.abs(inv.cpu().numpy() - g[f"inv.{tag}"])
        if nearest:
            # The inverse samples the 1 mm grid at c = 1.25 * i: every fourth index per axis is an exact .5 tie, which the
            # reference (nearbyint on float32 coordinates that went through a normalise / un-normalise round trip) resolves by
            # round-off.  Away from ties the value must be the reference's; on a tie it must be one of the two neighbours.
            got, pr = inv.cpu().numpy(), g["pred"]
            lo, hi, tie_ax = [], [], []
            for n_out, n_in in zip(got.shape[1:], pr.shape[1:]):
                c = 1.25 * np.arange(n_out)
                tie = np.abs(c - np.floor(c) - 0.5) < 1e-6
                lo.append(np.clip(np.where(tie, np.floor(c), np.rint(c)).astype(int), 0, n_in - 1))
                hi.append(np.clip(np.where(tie, np.ceil(c), np.rint(c)).astype(int), 0, n_in - 1))
                tie_ax.append(tie)
            any_tie = tie_ax[0][:, None, None] | tie_ax[1][None, :, None] | tie_ax[2][None, None, :]
            assert (diff[:, ~any_tie] > 1e-4).mean() < 2e-3, tag
            ok = np.zeros(got.shape, dtype=bool)
            for sel in itertools.product((0, 1), repeat=3):
                idx = [(lo, hi)[s_][a] for a, s_ in enumerate(sel)]
                ok |= np.abs(pr[:, idx[0]][:, :, idx[1]][:, :, :, idx[2]] - got) < 1e-4
            assert ok.mean() > 1 - 2e-3, (tag, ok.mean())
        else:
            assert diff.max() < 2e-4, (tag, diff.max())
        np.testing.assert_allclose(np.asarray(inv.affine), g[f"inv.{tag}.affine"], atol=1e-6)
