"""Both DAMSM losses of one step from one fused native call (pretrain_damsm.py:128-129,
train.py:140-141).  Not part of the reference API: a convenience for callers that always want
``WordsLoss`` and ``SentenceLoss`` together; the results equal calling the two drop-ins."""
from __future__ import annotations

import dataclasses

from ..agb_native import native
from .damsm_core import DamsmConfig, damsm_losses, resolve_math, split_att_maps


class DAMSMLoss:
    def __init__(self, device, gamma1=4.0, gamma2=5.0, gamma3=10.0, wlambda=5.0, slambda=5.0, *, math="auto",
                 process_group=None, att_maps="list", eps=1e-8, max_words=None):
        self.device = device
        self.math = math                      # same choices as WordsLoss (losses/words_loss.py); "auto" resolves per call
        self.wcfg = DamsmConfig(float(gamma1), float(gamma2), float(gamma3), float(wlambda), float(eps),
                                native.AGB_MATH_FP32 if math == "auto" else native.MATH_NAMES[math], process_group,
                                want_att=att_maps is not None, max_words=max_words)
        self.scfg = DamsmConfig(gamma3=float(gamma3), lam=float(slambda), eps=float(eps), group=process_group)
        self.att_maps = att_maps

    def get_losses(self, img_features, cnn_code, words_emb, rnn_code, labels, cap_lens, class_ids):
        """Returns (wloss, sloss, att_maps)"""
        wcfg = self.wcfg
        if self.math == "auto":
            wcfg = dataclasses.replace(wcfg, math=resolve_math("auto", img_features, words_emb, wcfg.gamma1))
        wl, sl, att = damsm_losses(img_features, cnn_code, words_emb, rnn_code, labels, cap_lens, class_ids,
                                   wcfg, self.scfg)
        ih, iw = img_features.shape[2], img_features.shape[3]
        if self.att_maps == "list":
            maps = split_att_maps(att, cap_lens, ih, iw)
        elif self.att_maps == "packed":
            maps = att.reshape(att.shape[0], att.shape[1], ih, iw)
        else:
            maps = []
        return wl, sl, maps
