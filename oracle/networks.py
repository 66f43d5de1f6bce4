"""torch-CPU fp32 functional restatement of the reference networks on the hot path -- TEST INFRASTRUCTURE.

Each function rebuilds the forward pass from a reference `state_dict` with torch.nn.functional ops only (the same
ATen ops the reference modules dispatch to), citing the module it restates.  Pinned by tests/golden fixtures that
were produced by the real reference modules (tests/golden/make_golden.py).
"""
from __future__ import annotations

import itertools

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------- UNet
def _adn(x, sd, prefix, act_default=True):
    """ADN("NDA") with InstanceNorm(affine optional) + PReLU (monai/networks/blocks/acti_norm.py:19-101)."""
    w, b = sd.get(prefix + ".N.weight"), sd.get(prefix + ".N.bias")
    x = F.instance_norm(x, weight=w, bias=b, eps=1e-5)
    return F.prelu(x, sd[prefix + ".A.weight"])


def unet_forward(sd: dict, x: torch.Tensor, strides, prefix: str = "model", top: bool = True) -> torch.Tensor:
    """UNet with num_res_units=0 (monai/networks/nets/unet.py:150-182, 296-298): Sequential(down, Skip(sub), up)."""
    s = strides[0]
    k = sd[prefix + ".0.conv.weight"].shape[-1]
    pad = (k - 1) // 2
    d = F.conv3d(x, sd[prefix + ".0.conv.weight"], sd.get(prefix + ".0.conv.bias"), stride=s, padding=pad)
    d = _adn(d, sd, prefix + ".0.adn")
    sub = prefix + ".1.submodule"
    if (sub + ".conv.weight") in sd:  # bottom layer: a single stride-1 Convolution
        y = F.conv3d(d, sd[sub + ".conv.weight"], sd.get(sub + ".conv.bias"), stride=1, padding=pad)
        y = _adn(y, sd, sub + ".adn")
    else:
        y = unet_forward(sd, d, strides[1:], sub, top=False)
    c = torch.cat([d, y], dim=1)  # SkipConnection (layers/simplelayers.py:127-131)
    ku = sd[prefix + ".2.conv.weight"].shape[-1]
    u = F.conv_transpose3d(c, sd[prefix + ".2.conv.weight"], sd.get(prefix + ".2.conv.bias"), stride=s, padding=(ku - 1) // 2, output_padding=s - 1)
    if not top:
        u = _adn(u, sd, prefix + ".2.adn")
    return u


# ------------------------------------------------------------------------------------------------------ BasicUNet
def _two_conv(x, sd, prefix, slope=0.1):
    """TwoConv (monai/networks/nets/basic_unet.py:27-58): 2 x [conv3 pad1 -> InstanceNorm(affine) -> LeakyReLU(0.1)]."""
    for i in (0, 1):
        p = f"{prefix}.conv_{i}"
        x = F.conv3d(x, sd[p + ".conv.weight"], sd.get(p + ".conv.bias"), padding=1)
        x = F.instance_norm(x, weight=sd.get(p + ".adn.N.weight"), bias=sd.get(p + ".adn.N.bias"), eps=1e-5)
        x = F.leaky_relu(x, slope)
    return x


def basic_unet_forward(sd: dict, x: torch.Tensor) -> torch.Tensor:
    """BasicUNet default config (basic_unet.py:178-282): conv_0, down_1..4 (MaxPool2 + TwoConv), upcat_4..1, final 1x1."""
    x0 = _two_conv(x, sd, "conv_0")
    feats = [x0]
    h = x0
    for i in range(1, 5):
        h = _two_conv(F.max_pool3d(h, 2), sd, f"down_{i}.convs")
        feats.append(h)
    u = feats[4]
    for i, skip in zip((4, 3, 2, 1), (feats[3], feats[2], feats[1], feats[0])):
        p = f"upcat_{i}"
        u = F.conv_transpose3d(u, sd[p + ".upsample.deconv.weight"], sd.get(p + ".upsample.deconv.bias"), stride=2)
        # replicate-pad the upsampled tensor on the high side when the skip is larger (basic_unet.py:165-170)
        pads = []
        for d in range(u.dim() - 1, 1, -1):
            pads += [0, skip.shape[d] - u.shape[d]]
        if any(pads):
            u = F.pad(u, pads, "replicate")
        u = _two_conv(torch.cat([skip, u], dim=1), sd, p + ".convs")
    return F.conv3d(u, sd["final_conv.weight"], sd.get("final_conv.bias"))


# ------------------------------------------------------------------------------------------------------ SwinUNETR
def _res_block(x, sd, prefix):
    """UnetResBlock (monai/networks/blocks/dynunet_block.py:25-111), instance norm (non-affine), LeakyReLU(0.01)."""
    out = F.conv3d(x, sd[prefix + ".conv1.conv.weight"], padding=1)
    out = F.leaky_relu(F.instance_norm(out, eps=1e-5), 0.01)
    out = F.conv3d(out, sd[prefix + ".conv2.conv.weight"], padding=1)
    out = F.instance_norm(out, eps=1e-5)
    res = x
    if (prefix + ".conv3.conv.weight") in sd:
        res = F.instance_norm(F.conv3d(x, sd[prefix + ".conv3.conv.weight"]), eps=1e-5)
    return F.leaky_relu(out + res, 0.01)


def _up_block(x, skip, sd, prefix):
    """UnetrUpBlock (monai/networks/blocks/unetr_block.py:22-86): ConvTranspose k2 s2 -> cat(skip) -> UnetResBlock."""
    u = F.conv_transpose3d(x, sd[prefix + ".transp_conv.conv.weight"], stride=2)
    return _res_block(torch.cat([u, skip], dim=1), sd, prefix + ".conv_block")


def _window_partition(x, ws):
    b, d, h, w, c = x.shape
    x = x.view(b, d // ws[0], ws[0], h // ws[1], ws[1], w // ws[2], ws[2], c)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, ws[0] * ws[1] * ws[2], c)


def _window_reverse(win, ws, dims):
    b, d, h, w = dims
    x = win.view(b, d // ws[0], h // ws[1], w // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(b, d, h, w, -1)


def _compute_mask(dims, ws, ss):
    """compute_mask (monai/networks/nets/swin_unetr.py:779-816)."""
    d, h, w = dims
    img = torch.zeros((1, d, h, w, 1))
    cnt = 0
    for ds in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for hs in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for wsl in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                img[:, ds, hs, wsl, :] = cnt
                cnt += 1
    mw = _window_partition(img, ws).squeeze(-1)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


def _swin_block(x, sd, prefix, heads, window, shift, mask):
    """SwinTransformerBlock.forward (swin_unetr.py:596-698) with WindowAttention.forward (509-532)."""
    b, d, h, w, c = x.shape
    shortcut = x
    x = F.layer_norm(x, (c,), sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"], 1e-5)
    ws = [window[i] if (d, h, w)[i] > window[i] else (d, h, w)[i] for i in range(3)]
    ss = [shift[i] if (d, h, w)[i] > window[i] else 0 for i in range(3)]
    pd = (ws[0] - d % ws[0]) % ws[0]
    pb = (ws[1] - h % ws[1]) % ws[1]
    pr = (ws[2] - w % ws[2]) % ws[2]
    x = F.pad(x, (0, 0, 0, pr, 0, pb, 0, pd))
    _, dp, hp, wp, _ = x.shape
    if any(s > 0 for s in ss):
        x = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        am = mask
    else:
        am = None
    xw = _window_partition(x, ws)
    bw, n, _ = xw.shape
    qkv = F.linear(xw, sd[prefix + ".attn.qkv.weight"], sd.get(prefix + ".attn.qkv.bias"))
    qkv = qkv.reshape(bw, n, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (c // heads) ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = sd[prefix + ".attn.relative_position_index"][:n, :n].reshape(-1)
    bias = sd[prefix + ".attn.relative_position_bias_table"][idx].reshape(n, n, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if am is not None:
        nw = am.shape[0]
        attn = attn.view(bw // nw, nw, heads, n, n) + am.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    attn = attn.softmax(dim=-1)
    xo = (attn @ v).transpose(1, 2).reshape(bw, n, c)
    xo = F.linear(xo, sd[prefix + ".attn.proj.weight"], sd[prefix + ".attn.proj.bias"])
    x = _window_reverse(xo.view(-1, *ws, c), ws, [b, dp, hp, wp])
    if any(s > 0 for s in ss):
        x = torch.roll(x, shifts=(ss[0], ss[1], ss[2]), dims=(1, 2, 3))
    x = x[:, :d, :h, :w, :]
    x = shortcut + x
    y = F.layer_norm(x, (c,), sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"], 1e-5)
    y = F.linear(y, sd[prefix + ".mlp.linear1.weight"], sd[prefix + ".mlp.linear1.bias"])
    y = F.linear(F.gelu(y), sd[prefix + ".mlp.linear2.weight"], sd[prefix + ".mlp.linear2.bias"])
    return x + y


def _patch_merging(x, sd, prefix):
    """PatchMerging.forward (swin_unetr.py:749-773): the v0.9.0 slice order, LayerNorm(8C), Linear(8C->2C, no bias)."""
    b, d, h, w, c = x.shape
    x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2, 0, d % 2))
    parts = [x[:, i::2, j::2, k::2, :] for i, j, k in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))]
    x = torch.cat(parts, -1)
    x = F.layer_norm(x, (8 * c,), sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"], 1e-5)
    return F.linear(x, sd[prefix + ".reduction.weight"])


def swin_transformer_forward(sd, x, heads=(3, 6, 12, 24), window=(7, 7, 7), depths=(2, 2, 2, 2), prefix="swinViT"):
    """SwinTransformer.forward (swin_unetr.py:1055-1075) with normalize=True."""
    def proj_out(t):
        return F.layer_norm(t.permute(0, 2, 3, 4, 1), (t.shape[1],)).permute(0, 4, 1, 2, 3)

    x0 = F.conv3d(x, sd[prefix + ".patch_embed.proj.weight"], sd[prefix + ".patch_embed.proj.bias"], stride=2)
    outs = [proj_out(x0)]
    cur = x0
    shift = tuple(i // 2 for i in window)
    for li in range(4):
        lp = f"{prefix}.layers{li + 1}.0"
        t = cur.permute(0, 2, 3, 4, 1).contiguous()
        b, d, h, w, c = t.shape
        ws = [window[i] if (d, h, w)[i] > window[i] else (d, h, w)[i] for i in range(3)]
        ss = [shift[i] if (d, h, w)[i] > window[i] else 0 for i in range(3)]
        dims = [-(-d // ws[0]) * ws[0], -(-h // ws[1]) * ws[1], -(-w // ws[2]) * ws[2]]
        mask = _compute_mask(dims, ws, ss).to(device=t.device, dtype=t.dtype) if any(s > 0 for s in ss) else None
        for bi in range(depths[li]):
            t = _swin_block(t, sd, f"{lp}.blocks.{bi}", heads[li], window, (0, 0, 0) if bi % 2 == 0 else shift, mask)
        t = _patch_merging(t, sd, lp + ".downsample")
        cur = t.permute(0, 4, 1, 2, 3).contiguous()
        outs.append(proj_out(cur))
    return outs


def swin_unetr_forward(sd: dict, x: torch.Tensor, heads=(3, 6, 12, 24)) -> torch.Tensor:
    """SwinUNETR.forward (monai/networks/nets/swin_unetr.py:315-330)."""
    hs = swin_transformer_forward(sd, x, heads=heads)
    enc0 = _res_block(x, sd, "encoder1.layer")
    enc1 = _res_block(hs[0], sd, "encoder2.layer")
    enc2 = _res_block(hs[1], sd, "encoder3.layer")
    enc3 = _res_block(hs[2], sd, "encoder4.layer")
    dec4 = _res_block(hs[4], sd, "encoder10.layer")
    dec3 = _up_block(dec4, hs[3], sd, "decoder5")
    dec2 = _up_block(dec3, enc3, sd, "decoder4")
    dec1 = _up_block(dec2, enc2, sd, "decoder3")
    dec0 = _up_block(dec1, enc1, sd, "decoder2")
    out = _up_block(dec0, enc0, sd, "decoder1")
    return F.conv3d(out, sd["out.conv.conv.weight"], sd["out.conv.conv.bias"])


# -------------------------------------------------------------------------------------------------------- DynUNet
def _t3(v):
    return (int(v),) * 3 if isinstance(v, int) else tuple(int(i) for i in v)


def _dyn_pad(k, s):
    """get_padding / get_output_padding (monai/networks/blocks/dynunet_block.py:304-327)."""
    k, s = _t3(k), _t3(s)
    p = tuple(int((ki - si + 1) / 2) for ki, si in zip(k, s))
    return p, tuple(2 * pi + si - ki for pi, si, ki in zip(p, s, k))


def _dyn_norm(x, sd, prefix):
    return F.instance_norm(x, weight=sd.get(prefix + ".weight"), bias=sd.get(prefix + ".bias"), eps=1e-5)


def _dyn_block(x, sd, prefix, k, s, res):
    """UnetBasicBlock / UnetResBlock (dynunet_block.py:25-177)."""
    out = F.conv3d(x, sd[prefix + ".conv1.conv.weight"], None, stride=_t3(s), padding=_dyn_pad(k, s)[0])
    out = F.leaky_relu(_dyn_norm(out, sd, prefix + ".norm1"), 0.01)
    out = _dyn_norm(F.conv3d(out, sd[prefix + ".conv2.conv.weight"], None, stride=1, padding=_dyn_pad(k, 1)[0]), sd, prefix + ".norm2")
    if res:
        r = x
        if (prefix + ".conv3.conv.weight") in sd:
            r = _dyn_norm(F.conv3d(x, sd[prefix + ".conv3.conv.weight"], None, stride=_t3(s), padding=_dyn_pad(1, s)[0]), sd, prefix + ".norm3")
        out = out + r
    return F.leaky_relu(out, 0.01)


def dynunet_forward(sd: dict, x: torch.Tensor, kernel_size, strides, upsample_kernel_size, res_block: bool = False) -> torch.Tensor:
    """DynUNet in eval mode (monai/networks/nets/dynunet.py:265-272 with the skip chain of :24-53, 170-236): input block, down
    blocks, bottleneck, then transposed convolution + concat + basic block per level, 1x1 output block."""
    n = len(strides)
    skips = []
    cur = _dyn_block(x, sd, "input_block", kernel_size[0], strides[0], res_block)
    skips.append(cur)
    for i in range(n - 2):
        cur = _dyn_block(cur, sd, f"downsamples.{i}", kernel_size[1 + i], strides[1 + i], res_block)
        skips.append(cur)
    cur = _dyn_block(cur, sd, "bottleneck", kernel_size[-1], strides[-1], res_block)
    for j in range(n - 1):   # upsamples[j] pairs with level n - 2 - j
        lvl = n - 2 - j
        uk = upsample_kernel_size[::-1][j]
        pad, opad = _dyn_pad(uk, uk)
        p = f"upsamples.{j}"
        up = F.conv_transpose3d(cur, sd[p + ".transp_conv.conv.weight"], sd.get(p + ".transp_conv.conv.bias"), stride=_t3(uk), padding=pad, output_padding=opad)
        cur = _dyn_block(torch.cat((up, skips[lvl]), dim=1), sd, p + ".conv_block", kernel_size[1:][::-1][j], 1, False)
    return F.conv3d(cur, sd["output_block.conv.conv.weight"], sd["output_block.conv.conv.bias"])


# ------------------------------------------------------------------------------------------------------ SegResNet
def _seg_norm_act(x, sd, prefix, groups, slope):
    """get_norm_layer(("GROUP", {"num_groups": g})) | "instance", then ReLU / LeakyReLU (segresnet_block.py:60-66)."""
    if groups:
        x = F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-5)
    else:
        x = F.instance_norm(x, eps=1e-5)
    return F.relu(x) if slope is None else F.leaky_relu(x, slope)


def _seg_resblock(x, sd, prefix, groups, slope):
    """ResBlock.forward (monai/networks/blocks/segresnet_block.py:83-96)."""
    y = F.conv3d(_seg_norm_act(x, sd, prefix + ".norm1", groups, slope), sd[prefix + ".conv1.conv.weight"], None, padding=1)
    y = F.conv3d(_seg_norm_act(y, sd, prefix + ".norm2", groups, slope), sd[prefix + ".conv2.conv.weight"], None, padding=1)
    return y + x


def segresnet_forward(sd: dict, x: torch.Tensor, blocks_down=(1, 2, 2, 4), blocks_up=(1, 1, 1), groups: int = 8, slope=None,
                      upsample_mode: str = "nontrainable") -> torch.Tensor:
    """SegResNet.forward in eval mode (monai/networks/nets/segresnet.py:160-197): encode, reverse the skips, decode, conv_final."""
    x = F.conv3d(x, sd["convInit.conv.weight"], None, padding=1)
    down = []
    for i, n in enumerate(blocks_down):
        if i > 0:
            x = F.conv3d(x, sd[f"down_layers.{i}.0.conv.weight"], None, stride=2, padding=1)
        for j in range(n):
            x = _seg_resblock(x, sd, f"down_layers.{i}.{j + 1}", groups, slope)
        down.append(x)
    down.reverse()
    for i, n in enumerate(blocks_up):
        x = F.conv3d(x, sd[f"up_samples.{i}.0.conv.weight"], None)
        if upsample_mode == "deconv":
            x = F.conv_transpose3d(x, sd[f"up_samples.{i}.1.deconv.weight"], sd[f"up_samples.{i}.1.deconv.bias"], stride=2)
        else:
            x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
        x = x + down[i + 1]
        for j in range(n):
            x = _seg_resblock(x, sd, f"up_layers.{i}.{j}", groups, slope)
    x = _seg_norm_act(x, sd, "conv_final.0", groups, slope)
    return F.conv3d(x, sd["conv_final.2.conv.weight"], sd["conv_final.2.conv.bias"])


# ---------------------------------------------------------------------------------------------------------- UNETR
def _vit_forward(sd, x, heads, prefix="vit"):
    """ViT.forward with classification=False (monai/networks/nets/vit.py:118-133): conv patch embedding + position embeddings
    (blocks/patchembedding.py:172-184), TransformerBlocks (blocks/transformerblock.py:91-99, selfattention.py:170-217), final LayerNorm."""
    pe = prefix + ".patch_embedding"
    p = sd[pe + ".patch_embeddings.weight"].shape[2:]
    t = F.conv3d(x, sd[pe + ".patch_embeddings.weight"], sd[pe + ".patch_embeddings.bias"], stride=tuple(p))
    t = t.flatten(2).transpose(-1, -2) + sd[pe + ".position_embeddings"]
    hidden = []
    i = 0
    while f"{prefix}.blocks.{i}.norm1.weight" in sd:
        b = f"{prefix}.blocks.{i}"
        C = t.shape[-1]
        d = C // heads
        h = F.layer_norm(t, (C,), sd[b + ".norm1.weight"], sd[b + ".norm1.bias"], 1e-5)
        qkv = F.linear(h, sd[b + ".attn.qkv.weight"], sd.get(b + ".attn.qkv.bias"))
        B_, S_ = qkv.shape[:2]
        qkv = qkv.reshape(B_, S_, 3, heads, d).permute(2, 0, 3, 1, 4)          # "b h (qkv l d) -> qkv b l h d"
        att = torch.softmax(torch.einsum("blxd,blyd->blxy", qkv[0], qkv[1]) * d**-0.5, dim=-1)
        o = torch.einsum("bhxy,bhyd->bhxd", att, qkv[2]).permute(0, 2, 1, 3).reshape(B_, S_, C)   # "b l h d -> b h (l d)"
        t = t + F.linear(o, sd[b + ".attn.out_proj.weight"], sd[b + ".attn.out_proj.bias"])
        h = F.layer_norm(t, (C,), sd[b + ".norm2.weight"], sd[b + ".norm2.bias"], 1e-5)
        m = F.linear(F.gelu(F.linear(h, sd[b + ".mlp.linear1.weight"], sd[b + ".mlp.linear1.bias"])), sd[b + ".mlp.linear2.weight"], sd[b + ".mlp.linear2.bias"])
        t = t + m
        hidden.append(t)
        i += 1
    C = t.shape[-1]
    return F.layer_norm(t, (C,), sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"], 1e-5), hidden


def _unetr_norm(x, sd, prefix):
    if (prefix + ".running_mean") in sd:   # norm_name="batch" in eval mode
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd.get(prefix + ".weight"), sd.get(prefix + ".bias"), False, 0.0, 1e-5)
    return F.instance_norm(x, weight=sd.get(prefix + ".weight"), bias=sd.get(prefix + ".bias"), eps=1e-5)


def _unetr_conv_block(x, sd, prefix, res):
    """UnetResBlock / UnetBasicBlock with kernel 3, stride 1 (blocks/dynunet_block.py:25-177)."""
    out = F.leaky_relu(_unetr_norm(F.conv3d(x, sd[prefix + ".conv1.conv.weight"], None, padding=1), sd, prefix + ".norm1"), 0.01)
    out = _unetr_norm(F.conv3d(out, sd[prefix + ".conv2.conv.weight"], None, padding=1), sd, prefix + ".norm2")
    if res:
        r = x
        if (prefix + ".conv3.conv.weight") in sd:
            r = _unetr_norm(F.conv3d(x, sd[prefix + ".conv3.conv.weight"], None), sd, prefix + ".norm3")
        out = out + r
    return F.leaky_relu(out, 0.01)


def unetr_forward(sd: dict, x: torch.Tensor, num_heads: int, res_block: bool = True, conv_block: bool = True) -> torch.Tensor:
    """UNETR.forward (monai/networks/nets/unetr.py:196-213) with UnetrPrUpBlock / UnetrUpBlock / UnetrBasicBlock
    (blocks/unetr_block.py:22-259)."""
    tok, hidden = _vit_forward(sd, x, num_heads)
    p = sd["vit.patch_embedding.patch_embeddings.weight"].shape[2:]
    feat = tuple(s // k for s, k in zip(x.shape[2:], p))

    def proj_feat(t):
        return t.reshape(t.shape[0], *feat, t.shape[-1]).permute(0, 4, 1, 2, 3).contiguous()

    def pr_up(t, prefix):
        t = F.conv_transpose3d(t, sd[prefix + ".transp_conv_init.conv.weight"], None, stride=2)
        i = 0
        while True:
            if conv_block and f"{prefix}.blocks.{i}.0.conv.weight" in sd:
                t = F.conv_transpose3d(t, sd[f"{prefix}.blocks.{i}.0.conv.weight"], None, stride=2)
                t = _unetr_conv_block(t, sd, f"{prefix}.blocks.{i}.1", res_block)
            elif not conv_block and f"{prefix}.blocks.{i}.conv.weight" in sd:
                t = F.conv_transpose3d(t, sd[f"{prefix}.blocks.{i}.conv.weight"], None, stride=2)
            else:
                return t
            i += 1

    def up(t, skip, prefix):
        t = F.conv_transpose3d(t, sd[prefix + ".transp_conv.conv.weight"], None, stride=2)
        return _unetr_conv_block(torch.cat((t, skip), dim=1), sd, prefix + ".conv_block", res_block)

    enc1 = _unetr_conv_block(x, sd, "encoder1.layer", res_block)
    enc2 = pr_up(proj_feat(hidden[3]), "encoder2")
    enc3 = pr_up(proj_feat(hidden[6]), "encoder3")
    enc4 = pr_up(proj_feat(hidden[9]), "encoder4")
    dec3 = up(proj_feat(tok), enc4, "decoder5")
    dec2 = up(dec3, enc3, "decoder4")
    dec1 = up(dec2, enc2, "decoder3")
    out = up(dec1, enc1, "decoder2")
    return F.conv3d(out, sd["out.conv.conv.weight"], sd["out.conv.conv.bias"])
