"""model_lib.ControlNet.ldm.modules.diffusionmodules.openaimodel (reference: openaimodel.py:73-295,432-806)."""
from magicdance_b200.dropin.modules import (  # noqa: F401
    Downsample, ResBlock, TimestepBlock, TimestepEmbedSequential, UNetModel, Upsample)
