"""The DAMSM pretraining step around the native losses: seeded synthetic
"bedroom" captions with the reference's hierarchical-cluster structure, the trainable encoder heads with the
reference's state-dict names, and one optimiser step (losses -> backward -> clip -> Adam)."""
from .synthetic_captions import SyntheticBedroomCaptions, hierarchical_k_values  # noqa: F401
from .damsm_step import DamsmPretrainStep, RegionHeads, TextEncoder  # noqa: F401
