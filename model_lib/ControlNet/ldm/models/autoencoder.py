"""model_lib.ControlNet.ldm.models.autoencoder — the dotted path of the YAML's first_stage_config.target
(models/cldm_v15_reference_only_pose.yaml:75; reference: ldm/models/autoencoder.py:13-91).  Re-export only."""
from magicdance_b200.dropin.autoencoder import AutoencoderKL, DiagonalGaussianDistribution, IdentityFirstStage  # noqa: F401
