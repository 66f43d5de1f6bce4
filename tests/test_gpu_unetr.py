"""GPU parity of UNETR and its ViT kernels (monai/networks/nets/{unetr,vit}.py) against fixtures of the real reference."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from monai_b200 import _kernels as K
from monai_b200.inferers import sliding_window_inference
from monai_b200.networks.nets import UNETR
from oracle import networks as onet
from oracle import sliding_window as osw
from weights import fill_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cases(golden_dir):
    spec = importlib.util.spec_from_file_location("_unetr_cases", os.path.join(golden_dir, "unetr_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.UNETR_CASES


def _build(kw, seed):
    net = UNETR(**kw)
    net.load_state_dict(fill_state_dict(net.state_dict(), seed))
    return net.eval().to(DEV)


def test_layernorm_and_attention_kernels_vs_torch():
    g = torch.Generator().manual_seed(7)
    for (N, C, S, heads) in [(2, 96, 70, 4), (1, 768, 216, 12), (3, 64, 33, 8)]:
        x = (torch.randn((N, C, S), generator=g) * 1.5 + 0.2).to(DEV)
        gam, bet = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
        got = K.layernorm_cf(x, gam, bet, 1e-5)
        ref = F.layer_norm(x.transpose(1, 2), (C,), gam, bet, 1e-5).transpose(1, 2)
        assert float((got - ref).abs().max()) < 2e-5
        d = C // heads
        qkv = torch.randn((N, 3 * C, S), generator=g).to(DEV)
        got = K.mhsa_cf(qkv, heads, d, d**-0.5)
        q, k, v = qkv.reshape(N, 3, heads, d, S).unbind(1)                       # [N, heads, d, S]
        att = torch.softmax(torch.einsum("nhdx,nhdy->nhxy", q, k) * d**-0.5, dim=-1)
        ref = torch.einsum("nhxy,nhdy->nhdx", att, v).reshape(N, C, S)
        assert float((got - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
        goth = K.mhsa_cf(qkv.half(), heads, d, d**-0.5)
        assert goth.dtype == torch.float16 and float((goth.float() - ref).abs().max()) < 2e-2


@pytest.mark.parametrize("i", [0, 1])
def test_unetr_matches_the_reference_fixture(golden_dir, i):
    g = np.load(os.path.join(golden_dir, "unetr.npz"))
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


def test_unetr_as_sliding_window_predictor_vs_oracle(golden_dir):
    kw, okw, _, seed = _cases(golden_dir)[0]
    net = _build(kw, seed)
    x = torch.randn(1, 1, 48, 40, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        y = sliding_window_inference(x.to(DEV), (32, 32, 32), 2, net, overlap=0.25, mode="gaussian")
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    ref = osw.sliding_window_inference(x.numpy(), (32, 32, 32), 2, lambda p: onet.unetr_forward(sd, torch.from_numpy(p), **okw).numpy(),
                                       overlap=0.25, mode="gaussian")
    err = float(np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err < 1e-3, err
    with pytest.raises(ValueError):
        net(torch.zeros(1, 1, 16, 32, 32, device=DEV))
