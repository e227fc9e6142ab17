"""Drop-in for the reference ``losses/sentence_loss.py`` (ku222/Attention-GAN), B200-native.

``SentenceLoss(device, gamma3, slambda).get_loss(cnn_code, rnn_code, labels, class_ids, eps=1e-8)``
keeps the reference's signature and semantics (losses/sentence_loss.py:6-50): gamma3-scaled cosine
matrix, class-id mask, cross-entropy both ways, times lambda.
"""
from __future__ import annotations

from .damsm_core import DamsmConfig, sentence_loss


class SentenceLoss:
    def __init__(self, device, gamma3=10.0, slambda=5.0, *, process_group=None):
        self.device = device
        self.gamma3 = gamma3
        self.slambda = slambda
        self.process_group = process_group

    def get_loss(self, cnn_code, rnn_code, labels, class_ids, eps=1e-8):
        cfg = DamsmConfig(gamma3=float(self.gamma3), lam=float(self.slambda), eps=float(eps),
                          group=self.process_group)
        return sentence_loss(cnn_code, rnn_code, labels, class_ids, cfg)
