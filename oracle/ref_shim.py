"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (read-only tree at /root/reference).

Only usable in the build container (the GPU box has no /root/reference).  It exists to
(1) pin oracle/restatement.py against the reference's own code and (2) generate the
golden fixtures under tests/golden/ (see oracle/make_golden.py).

The reference cannot be imported as shipped: it needs pytorch_lightning, omegaconf,
xformers, diffusers and clip (ldm/models/diffusion/ddpm.py:12,21;
ldm/modules/diffusionmodules/model.py:12-13; ldm/modules/motion_module.py:10-14;
ldm/modules/encoders/modules.py:7).  We register inert stand-ins for those modules,
import ldm.modules.attention FIRST so XFORMERS_IS_AVAILBLE is False and the vanilla
CrossAttention branch (attention.py:168-199) is the one that runs, and only then stub
xformers for diffusionmodules/model.py's unconditional import.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn
import yaml

REFERENCE_ROOT = os.environ.get("MAGICDANCE_REFERENCE_ROOT", "/root/reference")
YAML_REL = "model_lib/ControlNet/models/cldm_v15_reference_only_pose.yaml"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, YAML_REL))


class _AttrDict(dict):
    """dict with attribute access, enough for `config.model` / `config.get('params')`."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def _to_attr(o):
    if isinstance(o, dict):
        return _AttrDict({k: _to_attr(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_to_attr(v) for v in o]
    return o


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_mdb_stub", False):
        return

    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    class Callback:
        pass

    ident = lambda f: f
    pl = _mod("pytorch_lightning", LightningModule=LightningModule, Callback=Callback, _mdb_stub=True)
    util = _mod("pytorch_lightning.utilities")
    rz = _mod("pytorch_lightning.utilities.rank_zero", rank_zero_only=ident)
    dist = _mod("pytorch_lightning.utilities.distributed", rank_zero_only=ident)
    cb = _mod("pytorch_lightning.callbacks", Callback=Callback)
    pl.utilities, pl.callbacks = util, cb
    util.rank_zero, util.distributed = rz, dist

    class ListConfig(list):
        pass

    class OmegaConf:
        @staticmethod
        def load(path):
            with open(path) as f:
                return _to_attr(yaml.safe_load(f))

        @staticmethod
        def to_container(c, **k):
            return c

    oc = _mod("omegaconf", ListConfig=ListConfig, OmegaConf=OmegaConf)
    oc.listconfig = _mod("omegaconf.listconfig", ListConfig=ListConfig)

    class _Any(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    def register_to_config(f):
        return f

    df = _mod("diffusers")
    df.configuration_utils = _mod("diffusers.configuration_utils", ConfigMixin=object, register_to_config=register_to_config)
    df.modeling_utils = _mod("diffusers.modeling_utils", ModelMixin=nn.Module)
    df.utils = _mod("diffusers.utils", BaseOutput=object)
    df.utils.import_utils = _mod("diffusers.utils.import_utils", is_xformers_available=lambda: False)
    df.models = _mod("diffusers.models")
    df.models.attention = _mod("diffusers.models.attention", CrossAttention=_Any, FeedForward=_Any)
    _mod("clip")


_LOADED = {}


def load_reference():
    """Import the reference hot-path modules; returns a namespace of the classes used."""
    if _LOADED:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    # the reference imports itself as `model_lib.ControlNet...`; make sure ITS tree wins and
    # that no other `model_lib` (e.g. this repo's drop-in) is already imported.
    for k in list(sys.modules):
        if k == "model_lib" or k.startswith("model_lib."):
            raise RuntimeError("a different `model_lib` is already imported in this process; "
                               "run the reference shim in its own process")
    # The reference's `model_lib` is a namespace package (no __init__.py); a regular package of the same name
    # anywhere on sys.path — this repo's drop-in tree — would win regardless of order.  Hide every such entry
    # while `model_lib` is first bound; later submodule imports go through the bound package's __path__.
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [
        e for e in saved_path
        if e != REFERENCE_ROOT and not os.path.isfile(os.path.join(e or os.getcwd(), "model_lib", "__init__.py"))]
    _install_stubs()
    import importlib

    try:
        importlib.import_module("model_lib.ControlNet")
    finally:
        sys.path[:] = [REFERENCE_ROOT] + [e for e in saved_path if e != REFERENCE_ROOT]
    attention = importlib.import_module("model_lib.ControlNet.ldm.modules.attention")
    assert attention.XFORMERS_IS_AVAILBLE is False
    xf = _mod("xformers")
    xf.ops = _mod("xformers.ops")
    cldm = importlib.import_module("model_lib.ControlNet.cldm.cldm")
    ddim = importlib.import_module("model_lib.ControlNet.ldm.models.diffusion.ddim")
    util = importlib.import_module("model_lib.ControlNet.ldm.util")
    dutil = importlib.import_module("model_lib.ControlNet.ldm.modules.diffusionmodules.util")
    _LOADED.update(attention=attention, cldm=cldm, ddim=ddim, util=util, dutil=dutil)
    return _LOADED


def load_yaml():
    with open(os.path.join(REFERENCE_ROOT, YAML_REL)) as f:
        return _to_attr(yaml.safe_load(f))


def build_reference_ldm(overrides: dict | None = None, with_vae: bool = False):
    """Instantiate the reference's ControlLDMReferenceOnlyPose from its own yaml
    (cldm/model.py:24-28 -> ldm/util.py:72-79) with the CLIP text encoder swapped for the
    reference's IdentityEncoder (encoders/modules.py:19-22) and, unless with_vae, the VAE
    swapped for the reference's IdentityFirstStage (both are off the hot path).

    `overrides` patches the three net configs (e.g. nothing for the real config)."""
    ref = load_reference()
    cfg = load_yaml()
    p = cfg["model"]["params"]
    p["cond_stage_config"] = _to_attr({"target": "model_lib.ControlNet.ldm.modules.encoders.modules.IdentityEncoder"})
    if not with_vae:
        p["first_stage_config"] = _to_attr({"target": "model_lib.ControlNet.ldm.models.autoencoder.IdentityFirstStage"})
    for key in ("appearance_control_stage_config", "pose_control_stage_config", "unet_config"):
        p[key]["params"]["use_checkpoint"] = False  # CheckpointFunction (util.py:118-187) is a no-op for no_grad fwd
        if overrides:
            p[key]["params"].update(overrides)
    model = ref["util"].instantiate_from_config(cfg["model"])
    model.eval()
    return model


def cpu_sampler(model):
    """DDIMSampler_ReferenceOnly with register_buffer overridden: the reference's version
    force-moves every buffer to 'cuda' (ddim.py:353-357)."""
    ref = load_reference()

    class _S(ref["ddim"].DDIMSampler_ReferenceOnly):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    return _S(model)
