"""Importable alias of the ``attention-gan_b200/`` tree (a hyphen is not a legal module name).

``import attention_gan_b200`` puts ``attention-gan_b200/`` on ``sys.path`` so the reference's own
top-level import paths resolve to the B200-native drop-ins:

    from networks.attention import AttentionModule, func_attention
    from losses.words_loss import WordsLoss
    from losses.sentence_loss import SentenceLoss
"""
import os
import sys

PACKAGE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "attention-gan_b200")
if PACKAGE_DIR not in sys.path:
    sys.path.insert(0, PACKAGE_DIR)

from agb_native import native, ops  # noqa: E402,F401
from networks.attention import AttentionModule, GlobalAttention, func_attention  # noqa: E402,F401
from losses.words_loss import WordsLoss  # noqa: E402,F401
from losses.sentence_loss import SentenceLoss  # noqa: E402,F401
from losses.damsm_loss import DAMSMLoss  # noqa: E402,F401
