"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header declares, the
drop-in modules keep the reference's interface and refuse CPU tensors, and the host-side sharding
logic (world_size 2, gloo) reproduces the single-process loss and gradients."""
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "attngan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import attention_gan_b200 as pkg
    lib = pkg.native.lib()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/attngan_b200.h but not exported"
        assert s in pkg.native.SIGNATURES, f"{s} has no ctypes signature"
    assert set(pkg.native.SIGNATURES) == set(syms)
    assert lib.agb_version() == 100
    # pure host queries work without a GPU
    assert lib.agb_contrastive_workspace_bytes(48) == 4 * 48 * 4
    assert lib.agb_damsm_supported(18, 256, 289, pkg.native.AGB_MATH_FP32) == 1
    assert lib.agb_damsm_supported(65, 256, 289, pkg.native.AGB_MATH_FP32) == 0
    assert lib.agb_damsm_workspace_bytes(48, 48, 18, 256, 289, 0) > 0
    assert lib.agb_word_attn_bwd_workspace_bytes(2, 32, 4096, 256, 18) > 0


def test_argument_errors_are_reported_without_a_gpu():
    import attention_gan_b200 as pkg
    lib = pkg.native.lib()
    rc = lib.agb_contrastive_fwd(None, 4, None, None, 10.0, 5.0, 0, 4, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.agb_last_error()
    rc = lib.agb_word_attn_fwd(None, None, 0, 0, 0, None, None, None, 0, None, None, 1, 1, 1, 1, 65, 0, 1, None)
    assert rc == -2 and b"T=65" in lib.agb_last_error()


def test_dropin_interface_matches_reference_signatures():
    import inspect
    import attention_gan_b200 as pkg
    m = pkg.AttentionModule(nc_in=32, emb_dim=256)
    assert list(m.state_dict().keys()) == ["conv1.weight"]            # Generator.pkl compatibility
    assert tuple(m.conv1.weight.shape) == (32, 256, 1, 1) and m.conv1.bias is None
    assert m.mask is None
    assert list(inspect.signature(m.forward).parameters) == ["images", "words", "scaled"]
    assert list(inspect.signature(pkg.func_attention).parameters) == ["query", "context", "gamma1", "scaled"]
    wl = pkg.WordsLoss(torch.device("cpu"))
    assert (wl.gamma1, wl.gamma2, wl.gamma3, wl.wlambda) == (4.0, 5.0, 10.0, 5.0)
    assert list(inspect.signature(wl.get_loss).parameters) == ["img_features", "words_emb", "labels", "cap_lens", "class_ids"]
    sl = pkg.SentenceLoss(torch.device("cpu"))
    assert (sl.gamma3, sl.slambda) == (10.0, 5.0)
    assert list(inspect.signature(sl.get_loss).parameters) == ["cnn_code", "rnn_code", "labels", "class_ids", "eps"]
    x1, x2 = torch.randn(5, 7), torch.randn(5, 7)
    torch.testing.assert_close(wl.cosine_similarity(x1, x2), torch.nn.functional.cosine_similarity(x1, x2))
    # reference top-level import paths resolve to the drop-ins
    from networks.attention import AttentionModule
    from losses.words_loss import WordsLoss
    from losses.sentence_loss import SentenceLoss
    assert AttentionModule is pkg.AttentionModule and WordsLoss is pkg.WordsLoss and SentenceLoss is pkg.SentenceLoss


def test_no_cpu_fallback():
    import attention_gan_b200 as pkg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.func_attention(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.WordsLoss("cpu").get_loss(torch.zeros(2, 4, 2, 2), torch.zeros(2, 4, 3), torch.arange(2),
                                      torch.tensor([3, 3]), None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.SentenceLoss("cpu").get_loss(torch.zeros(2, 4), torch.zeros(2, 4), torch.arange(2), None)


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "attention-gan_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.replace("oracle_ops", ""), f"{f} mentions the oracle"


# ---- world_size-2 gloo test of the sharded exchange ------------------------------------------
def _single_process_reference(seed):
    sys.path.insert(0, ROOT)
    from oracle import closed_form as cf
    from oracle import ref_port as rp
    img, wrd, cnn, rnn, labels, lens, cls = rp.synth_damsm(6, T=5, D=16, hw=3, seed=seed, n_classes=3)
    B, D = img.shape[:2]
    wl, _, dc, dw = cf.words_loss_fwd_bwd(img.numpy().reshape(B, D, -1), wrd.numpy(), labels.numpy(), lens.numpy(), cls)
    sl, _, dcnn, drnn = cf.sentence_loss_fwd_bwd(cnn.numpy(), rnn.numpy(), labels.numpy(), cls)
    return (img, wrd, cnn, rnn, labels, lens, cls), (wl, sl, dc.reshape(img.shape), dw, dcnn, drnn)


def _worker(rank, world, port, seed, fused, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import attention_gan_b200  # noqa: F401  (puts the drop-ins on sys.path)
        import oracle_ops
        from losses import damsm_core as core
        (img, wrd, cnn, rnn, labels, lens, cls), _ = _single_process_reference(seed)
        n = img.shape[0] // world
        sl = slice(rank * n, rank * n + n)
        im = img[sl].clone().requires_grad_(True)
        wd = wrd[sl].clone().requires_grad_(True)
        cn = cnn[sl].clone().requires_grad_(True)
        rn = rnn[sl].clone().requires_grad_(True)
        loc_labels = torch.arange(n)
        wcfg = core.DamsmConfig(group=dist.group.WORLD, ops=oracle_ops)
        scfg = core.DamsmConfig(group=dist.group.WORLD, ops=oracle_ops)
        if fused:
            wl, sls, att = core.damsm_losses(im, cn, wd, rn, loc_labels, lens[sl], cls[sl], wcfg, scfg)
        else:
            wl, att = core.words_loss(im, wd, loc_labels, lens[sl], cls[sl], wcfg)
            sls = core.sentence_loss(cn, rn, loc_labels, cls[sl], scfg)
        (wl + sls).backward()
        maps = core.split_att_maps(att, lens[sl], 3, 3)
        out[rank] = dict(wl=wl.item(), sl=sls.item(), dimg=im.grad.numpy(), dwords=wd.grad.numpy(),
                         dcnn=cn.grad.numpy(), drnn=rn.grad.numpy(), maps=[m.numpy() for m in maps])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fused", [False, True])
def test_sharded_losses_equal_single_process(fused):
    world, seed = 2, 11
    port = 29500 + (os.getpid() % 2000) + (7 if fused else 0)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, seed, fused, out), nprocs=world, join=True)
    (img, wrd, cnn, rnn, labels, lens, cls), (wl0, sl0, dc0, dw0, dcnn0, drnn0) = _single_process_reference(seed)
    n = img.shape[0] // world
    sys.path.insert(0, ROOT)
    from oracle import ref_port as rp
    _, ref_maps = rp.words_loss(img.double(), wrd.double(), labels, lens, cls)
    for r in range(world):
        o = out[r]
        sl = slice(r * n, r * n + n)
        assert abs(o["wl"] - wl0) < 1e-5 * abs(wl0) and abs(o["sl"] - sl0) < 1e-5 * abs(sl0)
        np.testing.assert_allclose(o["dimg"], dc0[sl], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o["dwords"], dw0[sl], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o["dcnn"], dcnn0[sl], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o["drnn"], drnn0[sl], rtol=1e-4, atol=1e-6)
        for j, m in enumerate(o["maps"]):
            np.testing.assert_allclose(m, ref_maps[r * n + j].numpy(), rtol=1e-5, atol=1e-7)
