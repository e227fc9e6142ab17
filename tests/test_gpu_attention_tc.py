"""GPU parity of the persistent tensor-core attention kernels (16-bit feature maps) against the fp64
closed form of the reference formulas (networks/attention.py:25-79 and its autograd).

The shapes are chosen to hit the paths the persistent kernels have and the small fixtures do not:
several tiles per CTA, tile ranges that cross one or several sample boundaries (operands rebuilt per
segment, per-sample partial-sum slots), ragged last tiles, the limits of the compiled word / channel
ranges, and a gradient that also arrives through the attention maps.  Tolerances as in
test_gpu_parity.py: 1e-3 of the tensor scale on the fp32-internal value plus one rounding of the output dtype.
"""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import ref_port as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def agb():
    import attention_gan_b200 as pkg
    assert pkg.native.lib().agb_version() >= 100
    return pkg


def rel_err(x, ref):
    x = x.detach().double().cpu().numpy()
    ref = np.asarray(ref, np.float64)
    return np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-30)


CASES = [
    # B, hw, T, C, dattn
    (40, 64, 18, 32, False),    # 1280 tiles over the persistent grids: multi-tile ranges + sample boundaries
    (600, 16, 18, 32, False),   # 2 tiles per sample: one CTA walks through several samples
    (5, 48, 32, 16, True),      # T = 32 (limit of the tensor-core backward), C = 16, gradient through the maps
    (3, 40, 7, 32, True),       # HW = 1600 is not a multiple of the 128-pixel tile; fewer than 8 words
    (2, 24, 64, 32, False),     # T = 64: long-caption kernels (bf16: tensor-core backward, one CTA per SM; fp16: CUDA cores)
    (3, 32, 40, 32, True),      # T = 40: long-caption backward with a ragged last chunk of 8 words + gradient through the maps
    (20, 16, 50, 16, False),    # T = 50, C = 16, two tiles per sample: segments of several samples per CTA
    (2, 64, 64, 32, True),      # T = 64 with d attn: 32 tiles per sample, many tiles per CTA
    (7, 32, 18, 64, False),     # C = 64: forward on tensor cores (1 CTA per SM), backward on CUDA cores
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,hw,T,C,with_dattn", CASES)
def test_persistent_attention_matches_closed_form(agb, dtype, B, hw, T, C, with_dattn):
    E = 64
    tol = 1e-3 + (2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11)
    images, words, weight, mask, _ = rp.synth_attention(B, C, E, T, hw, seed=7 * B + hw)
    images = images.to(dtype)                                     # the values the kernel sees
    g = torch.Generator().manual_seed(B)
    dctx = torch.randn(B, C, hw, hw, generator=g).to(dtype)
    dattn = torch.randn(B, T, hw, hw, generator=g).to(dtype) if with_dattn else None
    mod = agb.AttentionModule(C, E).cuda()
    with torch.no_grad():
        mod.conv1.weight.copy_(weight.cuda())
    im = images.cuda().requires_grad_(True)
    wd = words.cuda().requires_grad_(True)
    mod.apply_mask(mask.cuda())
    ctx, attn = mod(im, wd)
    assert ctx.dtype == dtype and attn.dtype == dtype
    loss = (ctx.float() * dctx.cuda().float()).sum()
    if with_dattn:
        loss = loss + (attn.float() * dattn.cuda().float()).sum()
    loss.backward()

    W = weight.reshape(C, E).numpy()
    h64 = images.double().numpy().reshape(B, C, -1)
    rc, ra, _ = cf.word_attention_fwd(h64, words.numpy(), W, mask.numpy(), True)
    assert rel_err(ctx, rc.reshape(B, C, hw, hw)) <= tol, "context"
    assert rel_err(attn, ra.reshape(B, T, hw, hw)) <= tol, "attn"
    # masked words get exactly zero attention, every pixel's map sums to one
    am = attn.detach().float().cpu().numpy()
    assert np.all(am[mask.numpy() == 0] == 0.0)
    np.testing.assert_allclose(am.sum(1), 1.0, atol=T * tol)

    da = dattn.double().numpy().reshape(B, T, -1) if with_dattn else None
    dh, dwords, dW = cf.word_attention_bwd(h64, words.numpy(), W, mask.numpy(),
                                           dctx.double().numpy().reshape(B, C, -1), da, True)
    assert rel_err(im.grad, dh.reshape(B, C, hw, hw)) <= tol, "dimages"
    # d(W.e) contracts attn / ds over the pixels on the tensor cores in the I/O precision
    gtol = 5e-3 if dtype == torch.bfloat16 else 1e-3
    assert rel_err(wd.grad, dwords) <= gtol, "dwords"
    assert rel_err(mod.conv1.weight.grad, dW.reshape(C, E, 1, 1)) <= gtol, "dweight"


def test_persistent_attention_is_deterministic_and_batch_split_invariant(agb):
    """Same inputs -> bit-identical outputs and gradients; forward results of a sample do not depend on
    which other samples share the launch (different tile ranges / CTAs)."""
    B, C, E, T, hw = 24, 32, 64, 18, 64
    images, words, weight, mask, _ = rp.synth_attention(B, C, E, T, hw, seed=3)
    images = images.to(torch.bfloat16)
    dctx = torch.randn(B, C, hw, hw, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).cuda()
    mod = agb.AttentionModule(C, E).cuda()
    with torch.no_grad():
        mod.conv1.weight.copy_(weight.cuda())

    def run(sl):
        im = images[sl].cuda().requires_grad_(True)
        wd = words[sl].cuda().requires_grad_(True)
        mod.apply_mask(mask[sl].cuda())
        mod.conv1.weight.grad = None
        ctx, attn = mod(im, wd)
        ctx.backward(dctx[sl])
        return ctx, attn, im.grad, wd.grad, mod.conv1.weight.grad.clone()

    a = run(slice(0, B))
    b = run(slice(0, B))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    c = run(slice(5, 16))
    assert torch.equal(c[0], a[0][5:16]) and torch.equal(c[1], a[1][5:16]) and torch.equal(c[2], a[2][5:16])


# ------------------------------------------------------------------------------------------------
# GenNextStage's concat epilogue -- context written straight into cat((h, ctx), 1)
# reference: networks/generator_submodules.py:113-116
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,want_attn", [(torch.float32, True), (torch.bfloat16, True), (torch.float16, False)])
@pytest.mark.parametrize("B,hw,T,C", [(3, 16, 18, 32), (4, 40, 7, 32), (2, 12, 5, 20)])
def test_forward_into_concat_buffer_matches_cat_of_oracle(agb, dtype, want_attn, B, hw, T, C):
    E = 64
    images, words, weight, mask, _ = rp.synth_attention(B, C, E, T, hw, seed=31 * B + hw)
    images = images.to(dtype)
    g = torch.Generator().manual_seed(5)
    dout = torch.randn(B, 2 * C, hw, hw, generator=g).to(dtype)
    mod = agb.AttentionModule(C, E).cuda()
    with torch.no_grad():
        mod.conv1.weight.copy_(weight.cuda())
    mod.apply_mask(mask.cuda())
    im = images.cuda().requires_grad_(True)
    wd = words.cuda().requires_grad_(True)
    buf = torch.full((B, 2 * C, hw, hw), float("nan"), dtype=dtype, device="cuda")
    out, attn = mod.forward_into(im, wd, buf, want_attn=want_attn)
    assert out.data_ptr() == buf.data_ptr() and (attn is None) == (not want_attn)
    out.backward(dout.cuda())

    h = images.double().numpy().reshape(B, C, -1)
    W2 = weight.reshape(C, E).double().numpy()
    rc, ra, _ = cf.word_attention_fwd(h, words.numpy(), W2, mask.numpy(), True)
    ref_out = np.concatenate([h, rc], axis=1).reshape(B, 2 * C, hw, hw)        # torch.cat((h_code, c_code), 1)
    d = dout.double().numpy().reshape(B, 2 * C, -1)
    dh, dwords, dW = cf.word_attention_bwd(h, words.numpy(), W2, mask.numpy(), d[:, C:], None, True)
    dh = dh + d[:, :C]                                                          # cat backward: identity half
    tol = 1e-5 if dtype == torch.float32 else 1e-3 + (2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11)
    assert torch.equal(out[:, :C], im.detach())
    assert rel_err(out, ref_out) < tol
    if want_attn:
        assert rel_err(attn, ra.reshape(B, T, hw, hw)) < tol
    gtol = 2e-5 if dtype == torch.float32 else tol * 2
    assert rel_err(im.grad, dh.reshape(im.shape)) < gtol
    assert rel_err(wd.grad, dwords) < gtol
    assert rel_err(mod.conv1.weight.grad.reshape(C, E), dW) < gtol
    # same numbers as the two-step route of the reference (attention, then torch.cat)
    im2 = images.cuda().requires_grad_(True)
    ctx2, _ = mod(im2, words.cuda())
    two_step = torch.cat((im2, ctx2), 1)
    assert torch.equal(two_step, out)
