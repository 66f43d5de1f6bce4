"""Pin the CPU oracle (oracle/) against (a) golden arrays transcribed from the reference's own unit tests and
(b) fixtures produced by the real reference (tests/golden/make_golden.py).  CPU only."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from oracle import networks as onet
from oracle import sliding_window as osw
from weights import fill_state_dict


def _npz(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_planner_matches_reference(golden_dir):
    g = _npz(golden_dir, "planner.npz")
    for i in range(int(g["n"])):
        image, roi, ov = tuple(g[f"c{i}.image"]), tuple(g[f"c{i}.roi"]), tuple(g[f"c{i}.overlap"])
        image_p = tuple(max(a, b) for a, b in zip(image, roi))
        interval = osw.get_scan_interval(image_p, roi, ov)
        assert interval == tuple(g[f"c{i}.interval"])
        sl = osw.dense_patch_slices(image_p, roi, interval)
        starts = np.array([[s.start for s in w] for w in sl])
        np.testing.assert_array_equal(starts, g[f"c{i}.starts"])


def test_importance_map_bit_exact(golden_dir):
    g = _npz(golden_dir, "planner.npz")
    for j in range(int(g["n_imp"])):
        sig = g[f"imp{j}.sigma"].tolist()
        sig = sig[0] if len(sig) == 1 else tuple(sig)
        m = osw.compute_importance_map(tuple(g[f"imp{j}.patch"]), str(g[f"imp{j}.mode"]), sig)
        np.testing.assert_array_equal(m, g[f"imp{j}.map"])


def _cheap_predictor_np(x):
    ramp = np.arange(x.shape[-1], dtype=x.dtype) * np.asarray(0.01, dtype=x.dtype)
    a = x.mean(axis=1, keepdims=True, dtype=x.dtype) * np.asarray(1.5, dtype=x.dtype) + ramp
    b = np.tanh(x[:, :1]) - np.asarray(0.25, dtype=x.dtype)
    return np.concatenate([a, b], axis=1)


def test_sliding_window_matches_reference(golden_dir):
    g = _npz(golden_dir, "sliding_window.npz")
    for name in g["names"]:
        ov = g[f"{name}.overlap"].tolist()
        ov = ov[0] if len(ov) == 1 else tuple(ov)
        y = osw.sliding_window_inference(
            g[f"{name}.x"], tuple(int(v) for v in g[f"{name}.roi"]), int(g[f"{name}.bs"]), _cheap_predictor_np, ov,
            str(g[f"{name}.mode"]), 0.125, str(g[f"{name}.pad"]), float(g[f"{name}.cval"]),
        )
        np.testing.assert_allclose(y, g[f"{name}.y"], rtol=1e-5, atol=1e-5, err_msg=f"case {name}")


def test_sliding_window_multi_resolution(golden_dir):
    g = _npz(golden_dir, "sliding_window.npz")

    def multi(x):
        t = torch.from_numpy(x)
        return {"1": x + 1.0, "2": (torch.nn.functional.avg_pool3d(t, 2) * 2.0).numpy(), "3": x[..., ::4, ::4, ::4] - 3.0}

    r = osw.sliding_window_inference(g["multi.x"], (16, 16, 16), 3, multi, 0.5, "gaussian")
    for k in ("1", "2", "3"):
        np.testing.assert_allclose(r[k], g[f"multi.y{k}"], rtol=1e-5, atol=1e-5)


SIGMA_CONSTANT = np.array(
    [[[[3.0000] * 7, [3.0000] * 7, [3.3333] * 7, [3.6667] * 7, [4.3333] * 7, [4.5000] * 7, [5.0000] * 7]]]
)
SIGMA_GAUSSIAN = np.array(
    [[[
        [3.0, 3.0, 3.0, 3.0, 3.0, 3.0, 3.0],
        [3.0, 3.0, 3.0, 3.0, 3.0, 3.0, 3.0],
        [3.3271625, 3.3271623, 3.3271623, 3.3271623, 3.3271623, 3.3271623, 3.3271625],
        [3.6728377, 3.6728377, 3.6728377, 3.6728377, 3.6728377, 3.6728377, 3.6728377],
        [4.3271623, 4.3271623, 4.3271627, 4.3271627, 4.3271627, 4.3271623, 4.3271623],
        [4.513757, 4.513757, 4.513757, 4.513757, 4.513757, 4.513757, 4.513757],
        [4.9999995, 5.0, 5.0, 5.0, 5.0, 5.0, 4.9999995],
    ]]]
)


class SigmaPred:
    """the stateful predictor of the reference's test_sigma (adds 2, 3, 4, ... to successive window batches)."""

    def __init__(self):
        self.add = 1

    def __call__(self, data):
        self.add += 1
        return data + self.add


def test_sigma_goldens_from_reference_unit_test():
    """golden arrays transcribed from tests/inferers/test_sliding_window_inference.py:158-241 (test_sigma)."""
    inputs = np.ones((1, 1, 7, 7), dtype=np.float32)
    r = osw.sliding_window_inference(inputs, (3, 3), 10, SigmaPred(), 0.5, "constant", 1.0, "constant", -1)
    np.testing.assert_allclose(r, SIGMA_CONSTANT, rtol=1e-4)
    r = osw.sliding_window_inference(inputs, (3, 3), 10, SigmaPred(), 0.5, "gaussian", 1.0, "constant", -1)
    np.testing.assert_allclose(r, SIGMA_GAUSSIAN, rtol=1e-4)
    r = osw.sliding_window_inference(inputs, (3, 3), 10, SigmaPred(), 0.5, "gaussian", [1.0, 1.0])
    np.testing.assert_allclose(r, SIGMA_GAUSSIAN, rtol=1e-4)


def test_cval_golden_from_reference_unit_test():
    """tests/inferers/test_sliding_window_inference.py:243-267 (test_cval): padded with -1, predictor x+1 -> summed."""
    inputs = np.ones((1, 1, 3, 3), dtype=np.float32)

    def compute(data):
        return data + data.sum()

    r = osw.sliding_window_inference(inputs, (5, 5), 10, compute, 0.5, "constant", 0.125, "constant", -1.0)
    expected = np.ones((1, 1, 3, 3)) * -6.0
    np.testing.assert_allclose(r, expected, rtol=1e-4)


def _my_state_dict(factory, seed):
    with contextlib.redirect_stdout(io.StringIO()):
        net = factory()
    return fill_state_dict(net.state_dict(), seed)


def test_unet_oracle_matches_reference_fixture(golden_dir):
    from monai_b200.networks.nets import UNet

    for name, seed, factory, strides in [
        ("unet_tiny.npz", 0, lambda: UNet(3, 1, 2, (4, 8, 16), (2, 2)), (2, 2)),
        ("unet_c2_32.npz", 1, lambda: UNet(3, 1, 2, (16, 32, 64, 128, 256), (2, 2, 2, 2)), (2, 2, 2, 2)),
    ]:
        g = _npz(golden_dir, name)
        sd = _my_state_dict(factory, seed)
        y = onet.unet_forward(sd, torch.from_numpy(g["x"]), strides)
        np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-4, atol=1e-5, err_msg=name)


def test_basic_unet_oracle_matches_reference_fixture(golden_dir):
    from monai_b200.networks.nets import BasicUNet

    g = _npz(golden_dir, "basic_unet_tiny.npz")
    sd = _my_state_dict(lambda: BasicUNet(3, 1, 2, features=(4, 4, 8, 8, 16, 4)), 3)
    y = onet.basic_unet_forward(sd, torch.from_numpy(g["x"]))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag", ["64", "96x64x64"])
def test_swin_unetr_oracle_matches_reference_fixture(golden_dir, tag):
    from monai_b200.networks.nets import SwinUNETR

    g = _npz(golden_dir, f"swin_unetr_fs48_{tag}.npz")
    sd = _my_state_dict(lambda: SwinUNETR(in_channels=1, out_channels=2, feature_size=48), 4)
    x = torch.from_numpy(g["x"].astype(np.float32))
    with torch.no_grad():
        hs = onet.swin_transformer_forward(sd, x)
        y = onet.swin_unetr_forward(sd, x)
    np.testing.assert_allclose(hs[0].numpy()[..., ::4, ::4, ::4], g["h0_sub"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(hs[1].numpy()[..., ::2, ::2, ::2], g["h1_sub"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(hs[2].numpy()[:, ::4], g["h2"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(hs[4].numpy()[:, ::8], g["h4"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(y.numpy()[..., ::4, ::4, ::4], g["y_sub"], rtol=1e-3, atol=1e-4)
    assert abs(float(y.double().mean()) - float(g["y_mean"])) < 1e-4


def test_state_dict_keys_and_shapes_match_the_reference(golden_dir):
    """tests/golden/state_dict_keys.json was dumped from the reference modules (make_golden side)."""
    import json

    from monai_b200.networks.nets import BasicUNet, SwinUNETR, UNet

    want = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    with contextlib.redirect_stdout(io.StringIO()):
        nets = {
            "swin_unetr_fs48": SwinUNETR(in_channels=1, out_channels=2, feature_size=48),
            "unet_c2": UNet(3, 1, 2, (16, 32, 64, 128, 256), (2, 2, 2, 2)),
            "unet_res": UNet(3, 2, 3, (4, 8, 8), (2, 1), num_res_units=2),
            "basic_unet": BasicUNet(),
        }
    for name, net in nets.items():
        got = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert got == want[name], name
    # legacy bundles pass img_size: accepted and ignored
    with contextlib.redirect_stdout(io.StringIO()):
        SwinUNETR(img_size=96, in_channels=1, out_channels=2, feature_size=48)


def _dynunet_cases():
    sys_path_golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util

    spec = importlib.util.spec_from_file_location("_mk_golden_cases", os.path.join(sys_path_golden, "dynunet_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DYNUNET_CASES


def test_dynunet_oracle_and_module_tree_match_the_reference(golden_dir):
    """f4 (other predictors): oracle.networks.dynunet_forward against outputs of the real reference DynUNet (tests/golden/dynunet.npz:
    isotropic default with affine instance norm, anisotropic residual variant, deep supervision in eval mode), and the product's
    DynUNet has the reference's state_dict keys and shapes (tests/golden/dynunet_state_dict_keys.json)."""
    import json

    from monai_b200.networks.nets import DynUNet
    from weights import fill_state_dict

    g = np.load(os.path.join(golden_dir, "dynunet.npz"))
    want = json.load(open(os.path.join(golden_dir, "dynunet_state_dict_keys.json")))
    for i, (kw, shape, seed) in enumerate(_dynunet_cases()):
        net = DynUNet(**kw)
        sd = net.state_dict()
        assert {k: list(v.shape) for k, v in sd.items()} == want[f"c{i}"], i
        # blocks are registered twice (flat containers + skip chain): load through the module so that aliases resolve as in the reference
        net.load_state_dict(fill_state_dict(sd, seed))
        sd = net.state_dict()
        y = onet.dynunet_forward(sd, torch.from_numpy(g[f"c{i}.x"]), kw["kernel_size"], kw["strides"], kw["upsample_kernel_size"], kw.get("res_block", False))
        assert tuple(y.shape) == tuple(g[f"c{i}.y"].shape)
        np.testing.assert_allclose(y.numpy(), g[f"c{i}.y"], rtol=1e-4, atol=1e-4)


def test_segresnet_oracle_and_module_tree_match_the_reference(golden_dir):
    """f4: oracle.networks.segresnet_forward against outputs of the real reference SegResNet (tests/golden/segresnet.npz: defaults with
    GroupNorm + trilinear upsampling, transposed-convolution variant, instance norm + LeakyReLU), and the product's SegResNet has
    the reference's state_dict keys and shapes."""
    import importlib.util
    import json

    from monai_b200.networks.nets import SegResNet

    spec = importlib.util.spec_from_file_location("_segresnet_cases", os.path.join(golden_dir, "segresnet_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(golden_dir, "segresnet.npz"))
    want = json.load(open(os.path.join(golden_dir, "segresnet_state_dict_keys.json")))
    for i, (kw, okw, shape, seed) in enumerate(mod.SEGRESNET_CASES):
        net = SegResNet(**kw)
        assert {k: list(v.shape) for k, v in net.state_dict().items()} == want[f"c{i}"], i
        net.load_state_dict(fill_state_dict(net.state_dict(), seed))
        y = onet.segresnet_forward(net.state_dict(), torch.from_numpy(g[f"c{i}.x"]), **okw)
        np.testing.assert_allclose(y.numpy(), g[f"c{i}.y"], rtol=1e-4, atol=1e-4)


def test_unetr_oracle_and_module_tree_match_the_reference(golden_dir):
    """f4: oracle.networks.unetr_forward (ViT encoder + UNETR decoder) against outputs of the real reference UNETR
    (tests/golden/unetr.npz), and the product's UNETR has the reference's state_dict keys and shapes -- including the cross-attention
    containers every reference TransformerBlock registers."""
    import importlib.util
    import json

    from monai_b200.networks.nets import UNETR

    spec = importlib.util.spec_from_file_location("_unetr_cases", os.path.join(golden_dir, "unetr_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(golden_dir, "unetr.npz"))
    want = json.load(open(os.path.join(golden_dir, "unetr_state_dict_keys.json")))
    for i, (kw, okw, shape, seed) in enumerate(mod.UNETR_CASES):
        net = UNETR(**kw)
        assert {k: list(v.shape) for k, v in net.state_dict().items()} == want[f"c{i}"], i
        net.load_state_dict(fill_state_dict(net.state_dict(), seed))
        y = onet.unetr_forward(net.state_dict(), torch.from_numpy(g[f"c{i}.x"]), **okw)
        np.testing.assert_allclose(y.numpy(), g[f"c{i}.y"], rtol=1e-4, atol=1e-4)
