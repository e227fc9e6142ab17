"""GPU parity of the tcgen05 region-feature head against the reference's own arithmetic: a bias-free 1x1
convolution (networks/cnn_encoder.py:56,101; utilities/layers.py:46-48) evaluated in fp64, and torch autograd of it."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(x, ref):
    x = x.detach().double().cpu()
    ref = ref.detach().double().cpu()
    return ((x - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,Cin,Cout,hw", [(5, 768, 256, 17), (3, 128, 128, 8), (2, 192, 256, 13), (17, 768, 256, 17)])
def test_region_head_matches_conv1x1(B, Cin, Cout, hw):
    import attention_gan_b200 as agb
    g = torch.Generator().manual_seed(B * 100 + hw)
    x = torch.randn(B, Cin, hw, hw, generator=g).relu()                 # Mixed_6e ends in ReLUs: non-negative features
    head = agb.RegionFeatureHead(Cout, Cin).cuda()
    assert list(head.state_dict().keys()) == ["emb_features.weight"]
    w = head.emb_features.weight.detach().cpu()
    dfeat = torch.randn(B, Cout, hw, hw, generator=g) * 1e-5            # loss gradients are tiny: bf16 range, not fp16
    xd = x.cuda().requires_grad_(True)
    feat = head(xd)
    feat.backward(dfeat.cuda())
    # fp64 reference through torch's own conv
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    ref = torch.nn.functional.conv2d(x64, w64)
    ref.backward(dfeat.double())
    assert feat.shape == (B, Cout, hw, hw) and feat.dtype == torch.float32
    assert _rel(feat, ref) < 2e-5, _rel(feat, ref)                      # split-precision forward (torch's TF32 conv: 5e-4)
    assert _rel(head.emb_features.weight.grad, w64.grad) < 5e-3        # bf16 operands in the backward
    assert _rel(xd.grad, x64.grad) < 5e-3
    # frozen trunk: no dx requested
    head.zero_grad()
    head(x.cuda()).backward(dfeat.cuda())
    assert _rel(head.emb_features.weight.grad, w64.grad) < 5e-3


def test_region_head_feeds_the_loss_like_the_fp32_convolution():
    """DAMSM loss on features from the native head vs. on features from torch's fp32 conv: within the 1e-4 loss bound"""
    import attention_gan_b200 as agb
    from oracle import ref_port as rp
    B = 24
    g = torch.Generator().manual_seed(4)
    m6e = (torch.randn(B, 768, 17, 17, generator=g) * 0.5).relu().cuda()
    _, wrd, cnn, rnn, labels, lens, cls = rp.synth_damsm(B, seed=12, n_classes=6)
    head = agb.RegionFeatureHead(256).cuda()
    loss = agb.WordsLoss("cuda", math="fp32", att_maps=None)
    wl_native, _ = loss.get_loss(head(m6e), wrd.cuda(), labels.cuda(), lens.cuda(), cls)
    torch.backends.cudnn.allow_tf32 = False                      # the reference arithmetic is fp32 (CPU) / fp32 conv
    wl_torch, _ = loss.get_loss(head.emb_features(m6e), wrd.cuda(), labels.cuda(), lens.cuda(), cls)
    torch.backends.cudnn.allow_tf32 = True
    assert abs(wl_native.item() - wl_torch.item()) <= 1e-4 * abs(wl_torch.item()), (wl_native.item(), wl_torch.item())


def test_region_head_rejects_unsupported_shapes():
    import attention_gan_b200 as agb
    from attention_gan_b200.agb_native import native
    head = agb.RegionFeatureHead(96, 100).cuda()
    with pytest.raises(native.NativeError):
        head(torch.zeros(1, 100, 4, 4, device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        agb.RegionFeatureHead(256)(torch.zeros(1, 768, 17, 17))
