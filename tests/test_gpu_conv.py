"""-m gpu: the conv / transposed-conv blocks of the image encoder / decoder (im2col / col2im +
tcgen05 GEMM + LN/SiLU row kernels) against torch's conv2d / conv_transpose2d in fp64 -- the
reference's Conv2dSamePad(k=4, s=2) -> ImgChLayerNorm -> SiLU and ConvTranspose2d(k=4, s=2, p=1)
stacks (networks.py:448-585, 771-810) -- values and every gradient."""
import pytest
import torch
import torch.nn.functional as F

import parity_cases as pc

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _ln_silu_ref(y, g, b):           # y NCHW fp64
    y = y.permute(0, 2, 3, 1)
    y = F.layer_norm(y, (y.shape[-1],), g, b, 1e-3)
    return F.silu(y)                  # NHWC


@pytest.mark.parametrize("n,H,W,Cin,Cout", [(3, 8, 8, 3, 8), (2, 16, 8, 8, 16), (5, 4, 4, 32, 64), (1, 64, 64, 3, 32)])
def test_conv_ln_silu_block(pkg, device, n, H, W, Cin, Cout):
    g_ = torch.Generator().manual_seed(n * 100 + H)
    x = torch.randn(n, H, W, Cin, generator=g_).to(device).requires_grad_(True)
    Wt = (torch.randn(Cout, Cin, 4, 4, generator=g_) / (16 * Cin) ** 0.5).to(device).requires_grad_(True)
    g = (1 + 0.1 * torch.randn(Cout, generator=g_)).to(device).requires_grad_(True)
    b = (0.1 * torch.randn(Cout, generator=g_)).to(device).requires_grad_(True)
    wgt = torch.randn(n, H // 2, W // 2, Cout, generator=g_).to(device)
    out = pkg.kernels.conv_ln_silu(x.reshape(n * H * W, Cin), (n, H, W), Wt, g, b)
    grads = torch.autograd.grad((out.reshape(n, H // 2, W // 2, Cout) * wgt).sum(), [x, Wt, g, b])
    xd, Wd, gd, bd = (t.detach().double().requires_grad_(True) for t in (x, Wt, g, b))
    y = F.conv2d(F.pad(xd.permute(0, 3, 1, 2), (1, 1, 1, 1)), Wd, None, 2)
    ref = _ln_silu_ref(y, gd, bd)
    rg = torch.autograd.grad((ref * wgt.double()).sum(), [xd, Wd, gd, bd])
    assert pc.rel(out.reshape(ref.shape), ref) < TOL
    for name, a, r in zip(("dx", "dW", "dg", "db"), grads, rg):
        assert pc.rel(a, r) < TOL, name


@pytest.mark.parametrize("n,h,w,Cin,Cout,norm", [(3, 4, 4, 16, 8, True), (2, 8, 4, 8, 3, False), (4, 4, 4, 64, 32, True),
                                                 (1, 32, 32, 32, 3, False)])
def test_deconv_block(pkg, device, n, h, w, Cin, Cout, norm):
    g_ = torch.Generator().manual_seed(n * 10 + h)
    x = torch.randn(n, h, w, Cin, generator=g_).to(device).requires_grad_(True)
    Wt = (torch.randn(Cin, Cout, 4, 4, generator=g_) / (4 * Cin) ** 0.5).to(device).requires_grad_(True)
    wgt = torch.randn(n, 2 * h, 2 * w, Cout, generator=g_).to(device)
    if norm:
        g = (1 + 0.1 * torch.randn(Cout, generator=g_)).to(device).requires_grad_(True)
        b = (0.1 * torch.randn(Cout, generator=g_)).to(device).requires_grad_(True)
        out = pkg.kernels.deconv_block(x.reshape(n * h * w, Cin), (n, h, w), Wt, g, b)
        leaves = [x, Wt, g, b]
    else:
        bias = (0.1 * torch.randn(Cout, generator=g_)).to(device).requires_grad_(True)
        out = pkg.kernels.deconv_block(x.reshape(n * h * w, Cin), (n, h, w), Wt, bias=bias, shift=0.5)
        leaves = [x, Wt, bias]
    grads = torch.autograd.grad((out.reshape(n, 2 * h, 2 * w, Cout) * wgt).sum(), leaves)
    dl = [t.detach().double().requires_grad_(True) for t in leaves]
    y = F.conv_transpose2d(dl[0].permute(0, 3, 1, 2), dl[1], None if norm else dl[2], 2, 1)
    ref = _ln_silu_ref(y, dl[2], dl[3]) if norm else y.permute(0, 2, 3, 1) + 0.5
    rg = torch.autograd.grad((ref * wgt.double()).sum(), dl)
    assert pc.rel(out.reshape(ref.shape), ref) < TOL
    for i, (a, r) in enumerate(zip(grads, rg)):
        assert pc.rel(a, r) < TOL, i


def test_conv_encoder_decoder_modules_match_torch_layers(pkg, device):
    """The modules' kernel path against their own nn.Sequential stacks run by torch (cuDNN, fp32):
    same parameters, same state_dict names; output layout / flatten order included."""
    torch.manual_seed(0)
    enc = pkg.networks.ConvEncoder((64, 64, 3), depth=8).to(device)
    dec = pkg.networks.ConvDecoder(96, (3, 64, 64), depth=8).to(device)
    with torch.no_grad():
        for p in list(enc.parameters()) + list(dec.parameters()):
            p.add_(0.05 * torch.randn_like(p))
    obs = torch.rand(2, 3, 64, 64, 3, device=device)
    feat = torch.randn(2, 3, 96, device=device)
    enc.requires_grad_(True)
    dec.requires_grad_(True)
    e = enc(obs)
    x = (obs - 0.5).reshape((-1, 64, 64, 3)).permute(0, 3, 1, 2)
    e_ref = enc.layers(x).reshape(2, 3, -1)
    assert e.shape == e_ref.shape == (2, 3, enc.outdim)
    assert pc.rel(e, e_ref) < 1e-4
    d = dec(feat)
    y = dec._linear_layer(feat).reshape(-1, 4, 4, dec._embed_size // 16)
    d_ref = dec.layers(y.permute(0, 3, 1, 2)).reshape(2, 3, 3, 64, 64).permute(0, 1, 3, 4, 2) + 0.5
    assert d.shape == d_ref.shape == (2, 3, 64, 64, 3)
    assert pc.rel(d, d_ref) < 1e-4
    w = torch.randn_like(d)
    ga = torch.autograd.grad((d * w).sum() + e.sum(), list(enc.parameters()) + list(dec.parameters()))
    gb = torch.autograd.grad((d_ref * w).sum() + e_ref.sum(), list(enc.parameters()) + list(dec.parameters()))
    for (k, _), a, b in zip(list(enc.named_parameters()) + list(dec.named_parameters()), ga, gb):
        assert pc.rel(a, b) < 2e-4, k
