"""Binding layer of the B200-native word-region attention path: ``native`` loads the C-ABI shared
library (include/attngan_b200.h), ``ops`` wraps its entry points for torch CUDA tensors."""
from . import native, ops  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .pipeline import HostPrefetcher  # noqa: F401
