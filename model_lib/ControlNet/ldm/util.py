"""model_lib.ControlNet.ldm.util (reference: ldm/util.py:72-87): the YAML target -> class factory."""
from magicdance_b200.dropin.util import get_obj_from_str, instantiate_from_config  # noqa: F401


def exists(x):
    return x is not None


def default(val, d):
    return val if val is not None else (d() if callable(d) else d)
