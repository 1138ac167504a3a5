"""model_lib.ControlNet.cldm.model — the import path the MagicPose scripts use (test_tiktok.py:39) for
create_model / load_state_dict / get_state_dict.  Interface of the reference's cldm/model.py:7-28; the checkpoint
readers here are this repo's own (magicdance_b200.dropin.util)."""
from magicdance_b200.dropin.util import (create_model, get_state_dict, instantiate_from_config,  # noqa: F401
                                         load_state_dict)
