"""model_lib.ControlNet.ldm.modules.attention (reference: attention.py:50-77,146-199,253-385)."""
from magicdance_b200.dropin.modules import (  # noqa: F401
    BasicTransformerBlock, CrossAttention, FeedForward, GEGLU, SpatialTransformer)
