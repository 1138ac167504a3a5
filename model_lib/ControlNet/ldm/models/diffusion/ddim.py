"""model_lib.ControlNet.ldm.models.diffusion.ddim (reference: ddim.py:346-730)."""
from magicdance_b200.dropin.ddim import DDIMSampler_ReferenceOnly  # noqa: F401
