"""GPU parity of DynUNet (monai/networks/nets/dynunet.py) against fixtures of the real reference and the oracle."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from monai_b200.inferers import sliding_window_inference
from monai_b200.networks.nets import DynUNet
from oracle import networks as onet
from oracle import sliding_window as osw
from weights import fill_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cases():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dynunet_cases.py")
    spec = importlib.util.spec_from_file_location("_dynunet_cases", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DYNUNET_CASES


def _build(kw, seed):
    net = DynUNet(**kw)
    net.load_state_dict(fill_state_dict(net.state_dict(), seed))
    return net.eval().to(DEV)


@pytest.mark.parametrize("i", [0, 1, 2])
def test_dynunet_matches_the_reference_fixture(golden_dir, i):
    """fp32: <= 1e-3 relative (north star's bar for fp32 conv paths); fp16 storage: 2e-2 of the output scale."""
    g = np.load(os.path.join(golden_dir, "dynunet.npz"))
    kw, _, seed = _cases()[i]
    net = _build(kw, seed)
    x, want = torch.from_numpy(g[f"c{i}.x"]).to(DEV), g[f"c{i}.y"]
    with torch.no_grad():
        y = net(x)
    assert y.dtype == torch.float32 and tuple(y.shape) == tuple(want.shape)
    err = float(np.abs(y.cpu().numpy() - want).max() / np.abs(want).max())
    assert err < 1e-3, (i, err)
    with torch.no_grad():
        yh = net.half()(x.half())
    assert yh.dtype == torch.float16
    errh = float(np.abs(yh.float().cpu().numpy() - want).max() / np.abs(want).max())
    assert errh < 2e-2, (i, errh)


def test_dynunet_as_sliding_window_predictor_vs_oracle():
    """The call a bundle makes: sliding_window_inference(volume, roi, sw_batch, DynUNet) -- against the CPU oracle of both."""
    kw, _, seed = _cases()[1]
    net = _build(kw, seed)
    x = torch.randn(1, 2, 24, 48, 40, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        y = sliding_window_inference(x.to(DEV), (16, 32, 24), 2, net, overlap=0.5, mode="gaussian")
    sd = {k: v.cpu() for k, v in net.state_dict().items()}

    def pred(p):
        return onet.dynunet_forward(sd, torch.from_numpy(p), kw["kernel_size"], kw["strides"], kw["upsample_kernel_size"], True).numpy()

    ref = osw.sliding_window_inference(x.numpy(), (16, 32, 24), 2, pred, overlap=0.5, mode="gaussian")
    err = float(np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err < 1e-3, err


def test_dynunet_training_mode_with_deep_supervision_is_rejected():
    kw, shape, seed = _cases()[2]
    net = _build(kw, seed).train()
    with pytest.raises(RuntimeError):
        net(torch.zeros(shape, device=DEV))
