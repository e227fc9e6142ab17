"""GPU tests of the DAMSM pretraining step: the native step (encoders -> native losses ->
hand-written backward -> clip -> Adam) against the same step with the oracle's losses and torch autograd
(reference pretrain_damsm.py:110-134), and the product's CUDA-graph helper."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_port as rp

pytestmark = pytest.mark.gpu


def _batch(B, seed):
    from attention_gan_b200.pretrain import SyntheticBedroomCaptions
    d = SyntheticBedroomCaptions(4 * B, seed=seed)
    caps, lens, cls = d.batch(0, B)
    g = torch.Generator().manual_seed(seed)
    m6e = torch.randn(B, 768, 17, 17, generator=g) * 0.5
    pooled = torch.randn(B, 2048, generator=g) * 0.5
    return d, caps.cuda(), lens.cuda(), cls, m6e.cuda(), pooled.cuda()


@pytest.mark.parametrize("math,gtol", [("fp32", 2e-4), ("auto", 1e-2)])
def test_pretrain_step_matches_oracle_step(math, gtol):
    """two optimiser steps on one batch: native step vs. (same encoders) + oracle losses + autograd.  Compared: the
    loss of both steps and the CLIPPED gradients of every parameter after the first backward (the parameters
    themselves are a poor yardstick: Adam's first update is lr * sign(g) even where g is rounding noise)."""
    from attention_gan_b200.pretrain import DamsmPretrainStep
    B = 16
    d, caps, lens, cls, m6e, pooled = _batch(B, seed=11)
    # fp32 arm: the fp32 head too (cuDNN, TF32 off), so the whole step is comparable at 2e-4; "auto": tcgen05 head
    st = DamsmPretrainStep(d.vocab_size, "cuda", math=math, seed=5, native_head=(math != "fp32"))
    torch.backends.cudnn.allow_tf32 = False
    st.rnn.dropout.p = 0.0                                  # dropout off (train mode stays: cuDNN LSTM backward needs it)
    rnn0, heads0 = copy.deepcopy(st.rnn), copy.deepcopy(st.heads)
    params0 = list(rnn0.parameters()) + list(heads0.parameters())
    opt0 = torch.optim.Adam(params0, lr=2e-3, betas=(0.5, 0.999))
    labels = torch.arange(B, device="cuda")
    cls_np = cls.numpy()
    cls_dev = cls.to("cuda", torch.int32)
    for it in range(2):
        loss = st.step(caps, lens, cls_dev, m6e, pooled, labels)
        # reference-style step (pretrain_damsm.py:118-133) with the oracle's losses
        feats, code = heads0(m6e, pooled)
        wemb, semb = rnn0(caps, lens)
        opt0.zero_grad()
        wl, _ = rp.words_loss(feats, wemb, labels, lens, cls_np)
        sl = rp.sentence_loss(code, semb, labels, cls_np)
        (wl + sl).backward()
        torch.nn.utils.clip_grad_norm_(rnn0.parameters(), 0.25)
        opt0.step()
        ref = float(wl + sl)
        assert abs(loss.item() - ref) <= (1e-4 if it == 0 else 2e-2) * abs(ref), (it, loss.item(), ref)
        if it == 0:
            for (n, p), q in zip(list(st.rnn.named_parameters()) + list(st.heads.named_parameters()), params0):
                err = (p.grad - q.grad).abs().max().item() / max(q.grad.abs().max().item(), 1e-12)
                assert err < gtol, (n, err)
    torch.backends.cudnn.allow_tf32 = True
    # the optimiser moved: same direction in both arms (updates are +-lr where the gradient is not noise)
    moved = [(p - q).abs().max().item() for p, q in zip(st.params, params0)]
    assert max(moved) <= 2.5 * 2e-3 * 2


def test_graphed_pretrain_step_equals_eager():
    """the whole step (cuDNN LSTM, heads, native losses fwd+bwd, clip, capturable Adam) replayed from ONE CUDA graph
    with refilled static inputs gives the eager numbers (agb_native/graph.py)"""
    import attention_gan_b200 as agb
    from attention_gan_b200.pretrain import DamsmPretrainStep
    B = 32
    d, caps, lens, cls, m6e, pooled = _batch(B, seed=21)
    labels = torch.arange(B, device="cuda")
    cls_dev = cls.to("cuda", torch.int32)
    eager = DamsmPretrainStep(d.vocab_size, "cuda", math="f16", seed=9)
    graphed = DamsmPretrainStep(d.vocab_size, "cuda", math="f16", seed=9)
    for s in (eager, graphed):
        s.rnn.dropout.p = 0.0
    static = [t.clone() for t in (caps, lens, cls_dev, m6e, pooled)]
    snap = copy.deepcopy((graphed.rnn.state_dict(), graphed.heads.state_dict(), graphed.optimizer.state_dict()))
    step = agb.GraphedStep(lambda: graphed.step(static[0], static[1], static[2], static[3], static[4], labels))
    # warm-up + capture advanced the weights: restore, then replay on two different batches
    graphed.rnn.load_state_dict(snap[0]); graphed.heads.load_state_dict(snap[1])
    for st_ in graphed.optimizer.state.values():
        for v in st_.values():
            if torch.is_tensor(v):
                v.zero_()
    for it, seed in enumerate((21, 22)):
        d2, caps2, lens2, cls2, m2, p2 = _batch(B, seed=seed)
        for dst, src in zip(static, (caps2, lens2, cls2.to("cuda", torch.int32), m2, p2)):
            dst.copy_(src)
        lg = step.replay().clone()
        le = eager.step(caps2, lens2, cls2.to("cuda", torch.int32), m2, p2, labels)
        # step 0: identical arithmetic on identical weights; step 1: the weights went through one Adam update, which
        # is +-lr wherever a gradient is rounding noise (cuDNN's LSTM backward is not bit-reproducible)
        assert abs(lg.item() - le.item()) <= (1e-5 if it == 0 else 2e-3) * abs(le.item()), (it, lg.item(), le.item())
    for p, q in zip(graphed.params, eager.params):
        assert (p - q).abs().max().item() <= 2 * 2 * 2e-3
