"""GPU parity of SegResNet (monai/networks/nets/segresnet.py) against fixtures of the real reference."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from monai_b200.inferers import sliding_window_inference
from monai_b200.networks.blocks import UpSample
from monai_b200.networks.blocks.acti_norm import norm_act_from_modules
from monai_b200.networks.nets import SegResNet
from oracle import networks as onet
from oracle import sliding_window as osw
from weights import fill_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cases(golden_dir):
    spec = importlib.util.spec_from_file_location("_segresnet_cases", os.path.join(golden_dir, "segresnet_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SEGRESNET_CASES


def _build(kw, seed):
    net = SegResNet(**kw)
    net.load_state_dict(fill_state_dict(net.state_dict(), seed))
    return net.eval().to(DEV)


def test_group_norm_and_trilinear_upsample_kernels_vs_torch():
    g = torch.Generator().manual_seed(2)
    x = (torch.randn((2, 16, 5, 6, 7), generator=g) * 2 + 0.5).to(DEV)
    gn = torch.nn.GroupNorm(4, 16).to(DEV)
    with torch.no_grad():
        gn.weight.copy_(torch.randn(16, generator=g)); gn.bias.copy_(torch.randn(16, generator=g))
    got = norm_act_from_modules(x, gn, torch.nn.ReLU())
    ref = F.relu(F.group_norm(x, 4, gn.weight, gn.bias, eps=gn.eps))
    assert float((got - ref).abs().max()) < 1e-4
    up = UpSample(3, 16, 16, scale_factor=2, mode="nontrainable", interp_mode="linear", align_corners=False)
    got_u = up(x)
    ref_u = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
    assert tuple(got_u.shape) == tuple(ref_u.shape) and float((got_u - ref_u).abs().max()) < 1e-5


@pytest.mark.parametrize("i", [0, 1, 2])
def test_segresnet_matches_the_reference_fixture(golden_dir, i):
    g = np.load(os.path.join(golden_dir, "segresnet.npz"))
    kw, _, _, seed = _cases(golden_dir)[i]
    net = _build(kw, seed)
    want = g[f"c{i}.y"]
    with torch.no_grad():
        y = net(torch.from_numpy(g[f"c{i}.x"]).to(DEV))
    assert y.dtype == torch.float32 and tuple(y.shape) == tuple(want.shape)
    err = float(np.abs(y.cpu().numpy() - want).max() / np.abs(want).max())
    assert err < 1e-3, (i, err)
    with torch.no_grad():
        yh = net.half()(torch.from_numpy(g[f"c{i}.x"]).to(DEV).half())
    errh = float(np.abs(yh.float().cpu().numpy() - want).max() / np.abs(want).max())
    assert yh.dtype == torch.float16 and errh < 3e-2, (i, errh)


def test_segresnet_as_sliding_window_predictor_vs_oracle(golden_dir):
    kw, okw, _, seed = _cases(golden_dir)[0]
    net = _build(kw, seed)
    x = torch.randn(1, 1, 24, 40, 32, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        y = sliding_window_inference(x.to(DEV), (16, 32, 24), 2, net, overlap=0.25, mode="gaussian")
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    ref = osw.sliding_window_inference(x.numpy(), (16, 32, 24), 2, lambda p: onet.segresnet_forward(sd, torch.from_numpy(p), **okw).numpy(),
                                       overlap=0.25, mode="gaussian")
    err = float(np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err < 1e-3, err
