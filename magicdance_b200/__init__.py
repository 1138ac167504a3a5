"""magicdance_b200 — B200 (sm_100a) kernels and host schedule for MagicPose's DDIM denoising hot path.

`ops` binds the C ABI (include/magicdance_b200.h); `engine` schedules the reference's three
networks over it.  Importing the package does not touch CUDA; using it without the compiled
library or without an sm_100 GPU raises (there is no fallback path).
"""
import json as _json
import os as _os


def _apply_switch_defaults():
    """magicdance_b200/switch_defaults.json (absent until a GPU run has validated something): {"MDB_...": "value"}
    defaults for the library's switches, applied with setdefault — an explicitly set environment variable always
    wins.  This is how a measured opt-in becomes the default: one reviewed line in that file, no code change."""
    path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "switch_defaults.json")
    if not _os.path.isfile(path):
        return {}
    with open(path) as f:
        cfg = {str(k): str(v) for k, v in _json.load(f).items() if str(k).startswith("MDB_")}
    for k, v in cfg.items():
        _os.environ.setdefault(k, v)
    return cfg


SWITCH_DEFAULTS = _apply_switch_defaults()

from . import ops  # noqa: E402,F401

__all__ = ["ops", "SWITCH_DEFAULTS"]
