"""A stand-in for ``agb_native.ops`` built on the oracle (numpy fp64), used ONLY by the CPU tests of
the host-side sharding logic (losses/damsm_core.py) under gloo.  Same call signatures as the
native wrappers; never imported by the product."""
import numpy as np
import torch

from oracle import closed_form as cf


def require_cuda(*tensors):
    return None


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def damsm_fwd(img, words, cap_lens, gamma1, gamma2, eps=1e-8, row_offset=0, want_att=True, math=0,
              cnn=None, rnn=None, keep_ws=False, save=False):
    c, w, lens = _np(img).astype(np.float64), _np(words).astype(np.float64), _np(cap_lens)
    m = cf.words_similarity_fwd(c, w, lens, gamma1, gamma2, eps)
    att = None
    if want_att:
        Bi, T, R = c.shape[0], w.shape[2], c.shape[2]
        att = np.zeros((Bi, T, R))
        for b in range(Bi):
            i = row_offset + b
            L = int(lens[i])
            _, beta, _ = cf._pair_block(c[b:b + 1], w[i, :, :L], gamma1)
            att[b, :L] = beta[0].T
        att = torch.from_numpy(att).float()
    scos = None
    if cnn is not None:
        scos = sent_cos_fwd(cnn, rnn, eps)
    if keep_ws:
        return torch.from_numpy(m).float(), att, scos, None
    return torch.from_numpy(m).float(), att, scos


def damsm_bwd(img, words, cap_lens, gamma1, gamma2, eps, dm, gscale, need_dwords, math=0, m_fwd=None, ws=None,
              ws_saved=False):
    g = 1.0 if gscale is None else float(gscale.item())
    dc, dw = cf.words_similarity_bwd(_np(img), _np(words), _np(cap_lens), _np(dm).astype(np.float64) * g,
                                     gamma1, gamma2, eps)
    dwords = torch.from_numpy(dw.transpose(0, 2, 1).copy()).float() if need_dwords else None
    return torch.from_numpy(dc).float(), dwords


def contrastive(raw, class_ids, labels, gamma3, lam, row_begin, row_count, want_grad=True):
    sim = gamma3 * _np(raw).astype(np.float64)
    cm = cf.class_mask(_np(class_ids)) if class_ids is not None else None
    if cm is not None:
        sim = np.where(cm, -np.inf, sim)
    loss, d = cf.two_way_ce_fwd_bwd(sim, _np(labels), lam=lam)
    if cm is not None:
        d = np.where(cm, 0.0, d)
    d = (d * gamma3)[row_begin:row_begin + row_count]
    return torch.tensor([loss], dtype=torch.float32), torch.from_numpy(d).float()


def sent_cos_fwd(cnn, rnn, eps):
    c, r = _np(cnn).astype(np.float64), _np(rnn).astype(np.float64)
    den = np.maximum(np.linalg.norm(c, axis=1)[:, None] * np.linalg.norm(r, axis=1)[None], eps)
    return torch.from_numpy(c @ r.T / den).float()


def sent_cos_bwd(cnn, rnn, eps, dscos, gscale, need_dcnn=True, need_drnn=True):
    c, r = _np(cnn).astype(np.float64), _np(rnn).astype(np.float64)
    g = _np(dscos).astype(np.float64) * (1.0 if gscale is None else float(gscale.item()))
    p = np.linalg.norm(c, axis=1)[:, None]
    q = np.linalg.norm(r, axis=1)[None]
    num = c @ r.T
    den = np.maximum(p * q, eps)
    gg = g / den
    gn = np.where(p * q > eps, gg * num, 0.0)
    dcnn = gg @ r - gn.sum(1, keepdims=True) / (p * p) * c
    drnn = gg.T @ c - gn.sum(0)[:, None] / (q.T * q.T) * r
    return torch.from_numpy(dcnn).float(), torch.from_numpy(drnn).float()
