"""Pin oracle/transforms.py against the reference: transcribed unit-test goldens + fixtures from the real reference."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import transforms as otr

IMG23 = [[[1, 1, 1], [2, 2, 2], [3, 3, 3]], [[4, 4, 4], [5, 5, 5], [6, 6, 6]]]
# tests/transforms/test_gaussian_smooth.py:25-86 (2-D image with 2 channels of 3x3)
GAUSS_GOLDEN = [
    (1.5, [[[0.59167546, 0.69312394, 0.59167546], [0.7956997, 0.93213004, 0.7956997], [0.7668002, 0.8982755, 0.7668002]],
           [[1.6105323, 1.8866735, 1.6105323], [1.9892492, 2.3303251, 1.9892492], [1.7856569, 2.091825, 1.7856569]]]),
    (0.5, [[[0.8424794, 0.99864554, 0.8424794], [1.678146, 1.9892154, 1.678146], [1.9889624, 2.3576462, 1.9889624]],
           [[2.966061, 3.5158648, 2.966061], [4.1953645, 4.973038, 4.1953645], [4.112544, 4.8748655, 4.1125436]]]),
    ([1.5, 0.5], [[[0.8542037, 1.0125432, 0.8542037], [1.1487541, 1.3616928, 1.1487541], [1.1070318, 1.3122368, 1.1070318]],
                  [[2.3251305, 2.756128, 2.3251305], [2.8718853, 3.4042323, 2.8718853], [2.5779586, 3.0558217, 2.5779586]]]),
]


@pytest.mark.parametrize("sigma,expected", GAUSS_GOLDEN)
def test_gaussian_smooth_reference_unit_test_goldens(sigma, expected):
    out = otr.gaussian_smooth(torch.tensor(IMG23, dtype=torch.float32), sigma)
    np.testing.assert_allclose(out.numpy(), np.array(expected), rtol=1e-4, atol=1e-4)


def _kw(g, tag):
    return ast.literal_eval(str(g[f"{tag}.kw"]))


def test_spacing_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    img = torch.from_numpy(g["img"])
    for tag in ("s0", "s1", "s2", "s3", "s4"):
        kw = _kw(g, tag)
        y, new_affine = otr.spacing(img, g[f"{tag}.affine"], g[f"{tag}.pixdim"], diagonal=kw.get("diagonal", False), mode=kw.get("mode", "bilinear"),
                                    padding_mode=kw.get("padding_mode", "border"), align_corners=kw.get("align_corners", False))
        assert tuple(y.shape) == g[f"{tag}.y"].shape, tag
        np.testing.assert_allclose(new_affine, g[f"{tag}.new_affine"], rtol=1e-10, atol=1e-10, err_msg=tag)
        if kw.get("mode") == "nearest":
            assert (y.numpy() != g[f"{tag}.y"]).mean() < 1e-3
        else:
            np.testing.assert_allclose(y.numpy(), g[f"{tag}.y"], rtol=1e-5, atol=1e-5, err_msg=tag)


def test_rand_affine_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    img = torch.from_numpy(g["img2"])
    for tag in ("r0", "r1", "r2"):
        kw = _kw(g, tag)
        y, _ = otr.rand_affine(img, 0, kw.get("rotate_range", ()), kw.get("shear_range", ()), kw.get("translate_range", ()), kw.get("scale_range", ()),
                               kw.get("spatial_size"), kw["mode"], kw["padding_mode"])
        assert tuple(y.shape) == g[f"{tag}.y"].shape, tag
        if kw["mode"] == "nearest":
            assert (y.numpy() != g[f"{tag}.y"]).mean() < 1e-3
        else:
            np.testing.assert_allclose(y.numpy(), g[f"{tag}.y"], rtol=1e-5, atol=1e-5, err_msg=tag)


def test_gaussian_smooth_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    img = torch.from_numpy(g["img2"])
    for tag in ("g0", "g1", "g2"):
        s = g[f"{tag}.sigma"].tolist()
        y = otr.gaussian_smooth(img, s[0] if len(s) == 1 else s)
        np.testing.assert_allclose(y.numpy(), g[f"{tag}.y"], rtol=1e-5, atol=1e-6)


def test_post_transforms_oracle_matches_reference_fixture(golden_dir):
    """Activations / AsDiscrete restatement vs outputs of the real reference (tests/golden/make_golden.py post), incl. argmax ties."""
    g = np.load(os.path.join(golden_dir, "post.npz"))
    logits = torch.from_numpy(g["logits"])
    np.testing.assert_allclose(otr.activations(logits, softmax=True).numpy(), g["softmax"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(otr.activations(logits, sigmoid=True).numpy(), g["sigmoid"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(otr.as_discrete(logits, argmax=True).numpy(), g["argmax"])
    np.testing.assert_array_equal(otr.as_discrete(logits, argmax=True, to_onehot=3).numpy(), g["argmax_onehot"])
    np.testing.assert_array_equal(otr.as_discrete(logits, threshold=0.25).numpy(), g["threshold"])
    r = torch.tensor([[0.5, 1.5, 2.5, -0.5, -1.5, 0.49, 2.51]])
    np.testing.assert_array_equal(otr.as_discrete(r, rounding="torchrounding").numpy(), g["round"])
    np.testing.assert_array_equal(otr.as_discrete(torch.from_numpy(g["labels"]), to_onehot=3).numpy(), g["onehot"])
    np.testing.assert_array_equal(otr.as_discrete(otr.activations(logits, sigmoid=True), threshold=0.5).numpy(), g["sigmoid_threshold"])
    assert g["argmax"][0, 0, 0, 0] == 0 and g["argmax"][0, 1, 1, 1] in (1.0, 2.0)
