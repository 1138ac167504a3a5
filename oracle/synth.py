"""TEST INFRASTRUCTURE — the synthetic weight / input generators live in magicdance_b200/synth.py
(bench.py and smoke() need them without touching oracle/); re-exported here for the oracle tools."""
from magicdance_b200.synth import *  # noqa: F401,F403
from magicdance_b200.synth import MANIFEST, SCHEDULE_KEYS, load_manifest, sample_indices, summarize, synth_inputs, synth_state_dict, synth_tensor  # noqa: F401
