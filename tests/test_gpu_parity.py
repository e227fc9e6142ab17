"""GPU parity tests: the native sm_100a kernels (through the C ABI and the drop-in modules) against
the golden vectors recorded from the unmodified reference and against the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): attention maps / contexts 1e-5 relative in fp32, 1e-3 in
reduced precision; loss values 1e-4 relative.  "Relative" is measured against the tensor's scale
(max |ref|) for tensors and against |ref| for scalars.  For bf16 storage the final store alone
rounds by up to 2^-8 relative (8 significand bits), so bf16 I/O is checked as: fp32-internal result within 1e-3, i.e.
output within 1e-3 + one bf16 rounding of the reference.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import closed_form as cf
from oracle import ref_port as rp

pytestmark = pytest.mark.gpu

ATTN_CASES = ["attn_small_scaled", "attn_small_unscaled", "attn_cfg1_slice", "attn_odd"]
DAMSM_CASES = ["damsm_small", "damsm_small_cls", "damsm_real_cls", "damsm_real_trained", "damsm_gammas"]


@pytest.fixture(scope="module")
def agb():
    import attention_gan_b200 as pkg
    assert pkg.native.lib().agb_version() >= 100
    return pkg


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def rel_err(x, ref):
    x = x.detach().double().cpu().numpy() if torch.is_tensor(x) else np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    return np.abs(x - ref).max() / scale


def assert_rel(x, ref, tol, what=""):
    e = rel_err(x, ref)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"


# ------------------------------------------------------------------------------------------------
# generator word attention
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ATTN_CASES)
@pytest.mark.parametrize("layout", ["transposed_view", "contiguous"])
def test_attention_module_fp32_matches_reference(agb, name, layout):
    g = load_golden(name)
    B, C, H, W = g["images"].shape
    E, T = g["words"].shape[1], g["words"].shape[2]
    mod = agb.AttentionModule(C, E).cuda()
    with torch.no_grad():
        mod.conv1.weight.copy_(dev(g["weight"]))
    im = dev(g["images"]).requires_grad_(True)
    if layout == "transposed_view":       # the RNN's physical layout (rnn_encoder.py:92)
        wd = dev(g["words"].transpose(0, 2, 1)).requires_grad_(True)
        words_in = wd.transpose(1, 2)
    else:
        wd = dev(g["words"]).requires_grad_(True)
        words_in = wd
    mod.apply_mask(dev(g["mask"], torch.int64))
    ctx, attn = mod(im, words_in, scaled=bool(g["scaled"]))
    assert ctx.shape == (B, C, H, W) and attn.shape == (B, T, H, W)
    assert_rel(ctx, g["ctx_f64"], 1e-5, "context")
    assert_rel(attn, g["attn_f64"], 1e-5, "attn")
    (ctx * dev(g["dctx"])).sum().add((attn * dev(g["dattn"])).sum()).backward()
    assert_rel(im.grad, g["dimages"], 2e-5, "dimages")
    dwords = wd.grad.transpose(1, 2) if layout == "transposed_view" else wd.grad
    assert_rel(dwords, g["dwords"], 2e-5, "dwords")
    assert_rel(mod.conv1.weight.grad, g["dweight"], 2e-5, "dweight")


def test_attention_module_without_dattn_and_frozen_words(agb):
    """train.py: attention maps never enter a loss and the RNN is frozen (train.py:89)"""
    g = load_golden("attn_cfg1_slice")
    B, C, H, W = g["images"].shape
    E = g["words"].shape[1]
    mod = agb.AttentionModule(C, E).cuda()
    with torch.no_grad():
        mod.conv1.weight.copy_(dev(g["weight"]))
    im = dev(g["images"]).requires_grad_(True)
    mod.apply_mask(dev(g["mask"], torch.int64))
    ctx, _ = mod(im, dev(g["words"]))
    (ctx * dev(g["dctx"])).sum().backward()
    dh, _, dW = cf.word_attention_bwd(g["images"].reshape(B, C, -1), g["words"], g["weight"].reshape(C, E),
                                      g["mask"], g["dctx"].reshape(B, C, -1), None, True)
    assert_rel(im.grad, dh.reshape(B, C, H, W), 2e-5, "dimages")
    assert_rel(mod.conv1.weight.grad, dW.reshape(C, E, 1, 1), 2e-5, "dweight")


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-3 + 2.0 ** -8), (torch.float16, 1e-3)])
def test_attention_module_reduced_precision_io(agb, dtype, tol):
    B, C, E, T, hw = 4, 32, 256, 18, 32
    images, words, weight, mask, _ = rp.synth_attention(B, C, E, T, hw, seed=21)
    images = images.to(dtype)                                     # the values the kernel sees
    g = torch.Generator().manual_seed(5)
    dctx = torch.randn(B, C, hw, hw, generator=g).to(dtype)
    mod = agb.AttentionModule(C, E).cuda()
    with torch.no_grad():
        mod.conv1.weight.copy_(weight.cuda())
    im = images.cuda().requires_grad_(True)
    wd = words.cuda().requires_grad_(True)
    mod.apply_mask(mask.cuda())
    ctx, attn = mod(im, wd)
    assert ctx.dtype == dtype and attn.dtype == dtype
    (ctx.float() * dctx.cuda().float()).sum().backward()
    h64 = images.double().numpy().reshape(B, C, -1)
    rc, ra, _ = cf.word_attention_fwd(h64, words.numpy(), weight.reshape(C, E).numpy(), mask.numpy(), True)
    assert_rel(ctx, rc.reshape(B, C, hw, hw), tol, "context")
    assert_rel(attn, ra.reshape(B, T, hw, hw), tol, "attn")
    dh, dwords, dW = cf.word_attention_bwd(h64, words.numpy(), weight.reshape(C, E).numpy(), mask.numpy(),
                                           dctx.double().numpy().reshape(B, C, -1), None, True)
    assert_rel(im.grad, dh.reshape(B, C, hw, hw), tol, "dimages")
    # d(W.e) contracts attn / ds over the pixels on the tensor cores in the I/O precision
    gtol = 5e-3 if dtype == torch.bfloat16 else 1e-3
    assert_rel(wd.grad, dwords, gtol, "dwords")
    assert_rel(mod.conv1.weight.grad, dW.reshape(C, E, 1, 1), gtol, "dweight")


@pytest.mark.parametrize("T,hw,C", [(1, 5, 4), (64, 9, 32), (33, 8, 64), (18, 3, 1)])
def test_attention_edge_shapes(agb, T, hw, C):
    """T = 1 / 64 (compiled limits), C = 1 / 64, ragged pixel counts (no vector alignment)"""
    B, E = 3, 24
    images, words, weight, mask, _ = rp.synth_attention(B, C, E, T, hw, seed=T + hw)
    mod = agb.AttentionModule(C, E).cuda()
    with torch.no_grad():
        mod.conv1.weight.copy_(weight.cuda())
    mod.apply_mask(mask.cuda())
    ctx, attn = mod(images.cuda(), words.cuda())
    rc, ra, _ = cf.word_attention_fwd(images.numpy().reshape(B, C, -1), words.numpy(), weight.reshape(C, E).numpy(),
                                      mask.numpy(), True)
    assert_rel(ctx, rc.reshape(ctx.shape), 1e-5, "context")
    assert_rel(attn, ra.reshape(attn.shape), 1e-5, "attn")
    # masked words get exactly zero attention
    am = attn.detach().cpu().numpy()
    assert np.all(am[mask.numpy() == 0] == 0.0)
    np.testing.assert_allclose(am.sum(1), 1.0, rtol=1e-5)


def test_attention_all_masked_sample_is_nan_like_reference(agb):
    images, words, weight, mask, _ = rp.synth_attention(2, 8, 16, 5, 4, seed=3)
    mask[1] = 0
    mod = agb.AttentionModule(8, 16).cuda()
    mod.apply_mask(mask.cuda())
    ctx, attn = mod(images.cuda(), words.cuda())
    assert torch.isnan(ctx[1]).all() and torch.isnan(attn[1]).all()
    assert torch.isfinite(ctx[0]).all()


def test_attention_rejects_cpu_and_missing_mask(agb):
    mod = agb.AttentionModule(8, 16)
    with pytest.raises(AttributeError):
        mod(torch.zeros(1, 8, 2, 2), torch.zeros(1, 16, 3))
    mod.apply_mask(torch.ones(1, 3, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(torch.zeros(1, 8, 2, 2), torch.zeros(1, 16, 3))


def test_attention_full_size_properties(agb):
    """cfg1 / cfg3-sized run: rows of attn sum to 1, context lies in the span of W.e, and the
    result does not depend on how the batch is split (size-independent properties)."""
    B, C, E, T, hw = 16, 32, 256, 18, 64
    images, words, weight, mask, _ = rp.synth_attention(B, C, E, T, hw, seed=0)
    mod = agb.AttentionModule(C, E).cuda()
    with torch.no_grad():
        mod.conv1.weight.copy_(weight.cuda())
    im, wd, mk = images.cuda(), words.cuda(), mask.cuda()
    mod.apply_mask(mk)
    ctx, attn = mod(im, wd)
    torch.testing.assert_close(attn.sum(1), torch.ones_like(attn[:, 0]), rtol=1e-5, atol=1e-5)
    we = torch.einsum("ce,bet->bct", weight.reshape(C, E).cuda(), wd)
    torch.testing.assert_close(ctx, torch.einsum("bct,bthw->bchw", we, attn), rtol=1e-4, atol=1e-5)
    mod.apply_mask(mk[5:9])
    c2, a2 = mod(im[5:9].contiguous(), wd[5:9])
    assert torch.equal(c2, ctx[5:9]) and torch.equal(a2, attn[5:9])
    # against the oracle on a slice
    rc, ra, _ = cf.word_attention_fwd(images[:2].numpy().reshape(2, C, -1), words[:2].numpy(),
                                      weight.reshape(C, E).numpy(), mask[:2].numpy(), True)
    assert_rel(ctx[:2], rc.reshape(2, C, hw, hw), 1e-5, "context")
    assert_rel(attn[:2], ra.reshape(2, T, hw, hw), 1e-5, "attn")


# ------------------------------------------------------------------------------------------------
# functional attention (row a5)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["func_small", "func_real"])
def test_func_attention_matches_reference(agb, name):
    g = load_golden(name)
    q = dev(g["query"]).requires_grad_(True)
    c = dev(g["context"]).requires_grad_(True)
    wc, attn = agb.func_attention(q, c, gamma1=float(g["gamma1"]))
    assert wc.shape == g["wc"].shape and attn.shape == g["attn"].shape
    assert_rel(wc, g["wc"], 1e-5, "weightedContext")
    assert_rel(attn, g["attn"], 1e-5, "attn")
    (wc * dev(g["dwc"])).sum().backward()
    assert_rel(q.grad, g["dquery"], 5e-5, "dquery")
    assert_rel(c.grad, g["dcontext"], 5e-5, "dcontext")


def test_func_attention_grad_through_attn_output(agb):
    g = torch.Generator().manual_seed(2)
    q = torch.randn(2, 16, 5, generator=g)
    c = torch.randn(2, 16, 3, 4, generator=g)
    da = torch.randn(2, 5, 3, 4, generator=g)
    dw = torch.randn(2, 16, 5, generator=g)
    q64 = q.double().requires_grad_(True)
    c64 = c.double().requires_grad_(True)
    wc, at = rp.func_attention(q64, c64, 3.0, scaled=False)
    ((wc * dw.double()).sum() + (at * da.double()).sum()).backward()
    qg = q.cuda().requires_grad_(True)
    cg = c.cuda().requires_grad_(True)
    wc2, at2 = agb.func_attention(qg, cg, 3.0, scaled=False)
    ((wc2 * dw.cuda()).sum() + (at2 * da.cuda()).sum()).backward()
    assert_rel(wc2, wc.detach().numpy(), 1e-5)
    assert_rel(at2, at.detach().numpy(), 1e-5)
    assert_rel(qg.grad, q64.grad.numpy(), 5e-5, "dquery")
    assert_rel(cg.grad, c64.grad.numpy(), 5e-5, "dcontext")


# ------------------------------------------------------------------------------------------------
# DAMSM words / sentence loss (rows a8-a10)
# ------------------------------------------------------------------------------------------------
def _damsm_args(g):
    cls = g["class_ids"] if bool(g["has_class_ids"]) else None
    g1, g2, g3 = (float(x) for x in g["gammas"])
    lw, ls = (float(x) for x in g["lambdas"])
    return cls, g1, g2, g3, lw, ls


def _check_damsm(agb, g, math, loss_tol, map_tol, grad_tol, fused):
    cls, g1, g2, g3, lw, ls = _damsm_args(g)
    B, D, ih, iw = g["img"].shape
    T = g["words"].shape[2]
    im = dev(g["img"]).requires_grad_(True)
    wt = dev(g["words"].transpose(0, 2, 1)).requires_grad_(True)      # physical [B,T,D] (rnn_encoder.py:92)
    cn = dev(g["cnn"]).requires_grad_(True)
    rn = dev(g["rnn"]).requires_grad_(True)
    labels = dev(g["labels"], torch.int64)
    lens = dev(g["cap_lens"], torch.int64)
    if fused:
        L = agb.DAMSMLoss("cuda", g1, g2, g3, lw, ls, math=math)
        wl, sl, maps = L.get_losses(im, cn, wt.transpose(1, 2), rn, labels, lens, cls)
    else:
        wl, maps = agb.WordsLoss("cuda", g1, g2, g3, lw, math=math).get_loss(im, wt.transpose(1, 2), labels, lens, cls)
        sl = agb.SentenceLoss("cuda", g3, ls).get_loss(cn, rn, labels, cls)
    assert wl.dim() == 0 and sl.dim() == 0
    assert abs(wl.item() - float(g["wloss_f64"])) <= loss_tol * abs(float(g["wloss_f64"]))
    assert abs(sl.item() - float(g["sloss_f64"])) <= 1e-4 * abs(float(g["sloss_f64"]))
    assert len(maps) == B
    for i, m in enumerate(maps):
        Li = int(g["cap_lens"][i])
        assert m.shape == (1, Li, ih, iw)
        assert_rel(m[0], g["att_maps"][i, :Li], map_tol, f"att_map[{i}]")
    (wl + sl).backward()
    assert_rel(im.grad, g["dimg"], grad_tol, "d img_features")
    assert_rel(wt.grad.transpose(1, 2), g["dwords"], grad_tol, "d words_emb")
    assert_rel(cn.grad, g["dcnn"], 1e-4, "d cnn_code")
    assert_rel(rn.grad, g["drnn"], 1e-4, "d rnn_code")
    # padded word slots get exactly zero gradient
    wg = wt.grad.detach().cpu().numpy()
    for i in range(B):
        assert np.all(wg[i, int(g["cap_lens"][i]):] == 0.0)


@pytest.mark.parametrize("name", DAMSM_CASES)
@pytest.mark.parametrize("fused", [False, True])
def test_damsm_fp32_matches_reference(agb, name, fused):
    _check_damsm(agb, load_golden(name), "fp32", 1e-5, 1e-5, 1e-4, fused)


def _oracle_damsm(img, wrd, cnn, rnn, labels, lens, cls, gam=(4.0, 5.0, 10.0), lam=(5.0, 5.0)):
    B, D = img.shape[:2]
    c = img.double().numpy().reshape(B, D, -1)
    w = wrd.double().numpy()
    wl, sim, dc, dw = cf.words_loss_fwd_bwd(c, w, labels.numpy(), lens.numpy(), cls, gam[0], gam[1], gam[2], lam[0])
    sl, _, dcnn, drnn = cf.sentence_loss_fwd_bwd(cnn.numpy(), rnn.numpy(), labels.numpy(), cls, gam[2], lam[1])
    return wl, sl, dc.reshape(img.shape), dw, dcnn, drnn


@pytest.mark.parametrize("B,T,full,ncls,trained", [(16, 18, False, 5, False), (12, 18, True, None, 0.12),
                                                   (9, 7, True, 4, False), (5, 30, False, None, False)])
def test_damsm_fp32_matches_oracle_seeded(agb, B, T, full, ncls, trained):
    img, wrd, cnn, rnn, labels, lens, cls = rp.synth_damsm(B, T=T, seed=100 + B, full_len=full, n_classes=ncls,
                                                           trained_like=trained)
    wl0, sl0, dc0, dw0, dcnn0, drnn0 = _oracle_damsm(img, wrd, cnn, rnn, labels, lens, cls)
    im = img.cuda().requires_grad_(True)
    wd = wrd.cuda().requires_grad_(True)          # transposed view of [B,T,D]
    cn = cnn.cuda().requires_grad_(True)
    rn = rnn.cuda().requires_grad_(True)
    wl, _ = agb.WordsLoss("cuda", math="fp32").get_loss(im, wd, labels.cuda(), lens.cuda(), cls)
    sl = agb.SentenceLoss("cuda").get_loss(cn, rn, labels.cuda(), cls)
    assert abs(wl.item() - wl0) <= 1e-5 * abs(wl0)
    assert abs(sl.item() - sl0) <= 1e-5 * abs(sl0)
    (wl + sl).backward()
    assert_rel(im.grad, dc0, 1e-4, "d img_features")
    assert_rel(wd.grad, dw0, 1e-4, "d words_emb")
    assert_rel(cn.grad, dcnn0, 1e-4)
    assert_rel(rn.grad, drnn0, 1e-4)


def test_damsm_frozen_text_encoder_skips_dwords(agb):
    """train.py:89: the RNN is frozen, words_emb does not require grad"""
    img, wrd, cnn, rnn, labels, lens, cls = rp.synth_damsm(6, T=9, D=64, hw=6, seed=4)
    _, _, dc0, _, _, _ = _oracle_damsm(img, wrd, cnn, rnn, labels, lens, cls)
    im = img.cuda().requires_grad_(True)
    wl, _ = agb.WordsLoss("cuda", math="fp32").get_loss(im, wrd.cuda(), labels.cuda(), lens.cuda(), cls)
    wl.backward()
    assert_rel(im.grad, dc0, 1e-4)


def test_damsm_row_blocks_equal_full_matrix(agb):
    """the sharding identity at cfg2 size: row blocks computed with row_offset reproduce the full
    similarity matrix and the matched-pair attention maps bit for bit"""
    from attention_gan_b200.agb_native import ops
    B = 48
    img, wrd, cnn, rnn, labels, lens, cls = rp.synth_damsm(B, seed=0)
    img3 = img.cuda().reshape(B, 256, -1).contiguous()
    w = wrd.cuda()
    l32 = lens.cuda().to(torch.int32)
    m, att, scos = ops.damsm_fwd(img3, w, l32, 4.0, 5.0, 1e-8, 0, True, 0, cnn.cuda(), rnn.cuda())
    for k in range(4):
        sl = slice(12 * k, 12 * k + 12)
        mk, ak, sk = ops.damsm_fwd(img3[sl].contiguous(), w, l32, 4.0, 5.0, 1e-8, 12 * k, True, 0,
                                   cnn.cuda()[sl].contiguous(), rnn.cuda())
        assert torch.equal(mk, m[sl]) and torch.equal(ak, att[sl]) and torch.equal(sk, scos[sl])
    # spot-check columns against the oracle
    ref = cf.words_similarity_fwd(img.numpy().reshape(B, 256, -1)[:4], wrd.numpy(), lens.numpy())
    assert_rel(m[:4], ref, 1e-5, "similarity rows")
    # attention maps are distributions over the regions for live words, zero for padded ones
    a = att.cpu().numpy()
    for i in range(B):
        np.testing.assert_allclose(a[i, : int(lens[i])].sum(-1), 1.0, rtol=1e-5)
        assert np.all(a[i, int(lens[i]):] == 0.0)


def test_contrastive_matches_oracle(agb):
    from attention_gan_b200.agb_native import ops
    rng = np.random.default_rng(0)
    B = 37
    raw = rng.normal(size=(B, B)).astype(np.float32)
    cls = rng.integers(0, 6, size=B).astype(np.int32)
    labels = rng.permutation(B).astype(np.int64)          # arbitrary targets, not just arange
    loss, draw = ops.contrastive(dev(raw), None, dev(labels, torch.int64), 10.0, 5.0, 3, 20)
    sim_nc = 10.0 * raw.astype(np.float64)
    l1, d1 = cf.two_way_ce_fwd_bwd(sim_nc, labels, lam=5.0)
    assert abs(loss.item() - l1) <= 1e-5 * abs(l1)
    assert_rel(draw, (d1 * 10.0)[3:23], 1e-5, "draw rows")
    labels2 = np.arange(B, dtype=np.int64)
    cm2 = cf.class_mask(cls)
    sim2 = np.where(cm2, -np.inf, sim_nc)
    l2, d2 = cf.two_way_ce_fwd_bwd(sim2, labels2, lam=5.0)
    loss, draw = ops.contrastive(dev(raw), dev(cls, torch.int32), dev(labels2, torch.int64), 10.0, 5.0, 0, B)
    assert abs(loss.item() - l2) <= 1e-5 * abs(l2)
    assert_rel(draw, np.where(cm2, 0.0, d2) * 10.0, 1e-5, "draw masked")


def test_damsm_edge_cases(agb):
    """minimum caption length, every caption in one class, B = 1 is degenerate but finite"""
    img, wrd, cnn, rnn, labels, lens, _ = rp.synth_damsm(4, T=6, D=32, hw=3, seed=9)
    lens[:] = torch.tensor([1, 6, 2, 1])
    cls = np.zeros(4, np.int64)                             # all share a class: only the diagonal is finite
    wl0, sl0, dc0, dw0, _, _ = _oracle_damsm(img, wrd, cnn, rnn, labels, lens, cls)
    im = img.cuda().requires_grad_(True)
    wd = wrd.cuda().requires_grad_(True)
    wl, maps = agb.WordsLoss("cuda", math="fp32").get_loss(im, wd, labels.cuda(), lens.cuda(), cls)
    assert [m.shape[1] for m in maps] == [1, 6, 2, 1]
    assert abs(wl.item() - wl0) <= 1e-5 * max(abs(wl0), 1e-6) + 1e-7
    wl.backward()
    assert_rel(im.grad, dc0, 1e-4) if np.abs(dc0).max() > 0 else None
    img1, wrd1, _, _, labels1, lens1, _ = rp.synth_damsm(1, T=4, D=32, hw=3, seed=1)
    wl1, _ = agb.WordsLoss("cuda", math="fp32").get_loss(img1.cuda(), wrd1.cuda(), labels1.cuda(), lens1.cuda(), None)
    assert wl1.item() == 0.0                                # a 1x1 cross-entropy


def test_native_rejects_unsupported_shapes(agb):
    from attention_gan_b200.agb_native import native, ops
    with pytest.raises(native.NativeError):
        ops.damsm_fwd(torch.zeros(1, 32, 9, device="cuda"), torch.zeros(1, 32, 65, device="cuda"),
                      torch.ones(1, dtype=torch.int32, device="cuda"), 4.0, 5.0)
    mod = agb.AttentionModule(8, 16).cuda()
    mod.apply_mask(torch.ones(1, 65, dtype=torch.int64, device="cuda"))
    with pytest.raises(native.NativeError, match="T=65"):
        mod(torch.zeros(1, 8, 2, 2, device="cuda"), torch.zeros(1, 16, 65, device="cuda"))


@pytest.mark.parametrize("math,ragged", [("fp32", False), ("f16", False), ("f16x2", True), ("f16", True)])
def test_nccl_sharded_losses_equal_single_process(math, ragged):
    """2 ranks over NCCL: loss and all gradients equal the single-process result on the concatenated batch AND the
    fp64 oracle; `ragged`: the second rank's captions are padded to T - 1 only
    (needs >= 2 GPUs; the CPU suite covers the same logic over gloo)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "check_sharded.py"), "64", math] + \
        (["ragged"] if ragged else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
