"""model_lib.ControlNet.ldm.models.diffusion.ddpm (reference: ddpm.py:46-521,1313-1352,1803-2601)."""
from magicdance_b200.dropin.ddpm import DDPM, DiffusionWrapper, LatentDiffusionReferenceOnly  # noqa: F401
