"""Stage-by-stage comparison of monai_b200.SwinUNETR against the torch-CPU oracle (debug aid, GPU box)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from weights import fill_state_dict  # noqa: E402

from monai_b200 import _kernels as K  # noqa: E402
from monai_b200.networks.nets import SwinUNETR  # noqa: E402
from oracle import networks as onet  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu().numpy(), b.float().cpu().numpy()
    return float(np.abs(a - b).max() / max(1e-6, np.abs(b).max()))


def main():
    dev = torch.device("cuda")
    net = SwinUNETR(in_channels=1, out_channels=2, feature_size=48)
    net.load_state_dict(fill_state_dict(net.state_dict(), 4))
    net = net.eval().to(dev)
    sd = {k: (v.float() if v.is_floating_point() else v).cpu() for k, v in net.state_dict().items()}
    x = torch.randn(1, 1, 64, 64, 64, generator=torch.Generator().manual_seed(15)).half().float()
    with torch.no_grad():
        hs = onet.swin_transformer_forward(sd, x)
        x0 = F.conv3d(x, sd["swinViT.patch_embed.proj.weight"], sd["swinViT.patch_embed.proj.bias"], stride=2)
        xd = x.to(dev).half()
        vit = net.swinViT
        t0, _ = K.conv_cin1_nc8(xd, vit.patch_embed.proj.weight, vit.patch_embed.proj.bias, 2, 2, 0)
        print("patch_embed", rel(K.unpack_nc8(t0), x0))
        print("h0", rel(K.unpack_nc8(net._proj_out(t0)), hs[0]))
        # first block pieces
        blk = vit.layers1[0].blocks[0]
        from monai_b200.networks.nets.swin_unetr import _get_window_size
        ws, ss = _get_window_size(t0.sp, blk.window_size, blk.shift_size)
        src, region, nW, n = net._plan(t0.sp, ws, ss, dev)
        xw = K.layernorm_nc8(t0, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, src=src, out_sp=(1, nW, n))
        tt = x0.permute(0, 2, 3, 4, 1)
        tn = F.layer_norm(tt, (48,), sd["swinViT.layers1.0.blocks.0.norm1.weight"], sd["swinViT.layers1.0.blocks.0.norm1.bias"])
        pad = [(w - d % w) % w for d, w in zip(t0.sp, ws)]
        tn = F.pad(tn, (0, 0, 0, pad[2], 0, pad[1], 0, pad[0]))
        want = onet._window_partition(tn, ws)
        got = K.unpack_nc8(xw).reshape(1, 48, nW, n).permute(0, 2, 3, 1).reshape(-1, n, 48)
        print("ln+partition", rel(got, want))
        qkv, _ = K.gemm_tc(xw, net._wlin(blk.attn.qkv.weight, "dbg.qkv"), 48, 144, bias=blk.attn.qkv.bias)
        wq = F.linear(want, sd["swinViT.layers1.0.blocks.0.attn.qkv.weight"], sd["swinViT.layers1.0.blocks.0.attn.qkv.bias"])
        print("qkv gemm", rel(K.unpack_nc8(qkv).reshape(1, 144, nW, n).permute(0, 2, 3, 1).reshape(-1, n, 144), wq))
        t1 = net._swin_stage(t0, vit.layers1[0], "l1")
        print("h1", rel(K.unpack_nc8(net._proj_out(t1)), hs[1]))
        t2 = net._swin_stage(t1, vit.layers2[0], "l2")
        print("h2", rel(K.unpack_nc8(net._proj_out(t2)), hs[2]))
        t3 = net._swin_stage(t2, vit.layers3[0], "l3")
        print("h3", rel(K.unpack_nc8(net._proj_out(t3)), hs[3]))
        t4 = net._swin_stage(t3, vit.layers4[0], "l4")
        print("h4", rel(K.unpack_nc8(net._proj_out(t4)), hs[4]))
        enc0 = onet._res_block(x, sd, "encoder1.layer")
        e0 = net._res_block(None, 1, 0, net.encoder1.layer, "enc1", x_in_raw=xd)
        print("enc0", rel(K.unpack_nc8(e0), enc0))
        enc1 = onet._res_block(hs[0], sd, "encoder2.layer")
        e1 = net._res_block(K.pack_nc8(hs[0].to(dev).half()), 48, 0, net.encoder2.layer, "enc2")
        print("enc1 (from oracle h0)", rel(K.unpack_nc8(e1), enc1))
        dec4 = onet._res_block(hs[4], sd, "encoder10.layer")
        d4 = net._res_block(K.pack_nc8(hs[4].to(dev).half()), 768, 0, net.encoder10.layer, "enc10")
        print("dec4 (from oracle h4)", rel(K.unpack_nc8(d4), dec4))
        y = net(xd)
        ref = onet.swin_unetr_forward(sd, x)
        print("final", rel(y, ref), "argmax agree", float((y.float().cpu().argmax(1) == ref.argmax(1)).float().mean()))


if __name__ == "__main__":
    main()
