"""TEST INFRASTRUCTURE — generates the VAE fixtures by running the UNMODIFIED reference AutoencoderKL.

Run in the build container only (needs /root/reference):

    cd /tmp && python /root/repo/oracle/make_golden_vae.py

Writes
  magicdance_b200/vae_manifest.json  key -> shape of the reference's first_stage_model state_dict (248 tensors)
  tests/golden/vae16.npz             decode_first_stage of a 16x16 latent (B=2) -> 128x128 image, and the
                                     encoder moments of a 128x128 image (B=2); outputs stored whole,
                                     per-level activations as deterministic subsamples + moments
  tests/golden/vae64.npz             decode_first_stage at the headline size (64x64 latent -> 512x512, B=1),
                                     stored as subsample + moments
"""
from __future__ import annotations

import importlib
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO in sys.path:
    sys.path.remove(REPO)  # the repo's own drop-in `model_lib` must not shadow the reference's
sys.path.append(REPO)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


ref_shim = _load("mdb_oracle_ref_shim", os.path.join(REPO, "oracle", "ref_shim.py"))
synth = _load("mdb_oracle_synth", os.path.join(REPO, "oracle", "synth.py"))
vae_oracle = _load("mdb_oracle_vae", os.path.join(REPO, "oracle", "vae_restatement.py"))

GOLDEN = os.path.join(REPO, "tests", "golden")
MANIFEST = os.path.join(REPO, "magicdance_b200", "vae_manifest.json")
SEED = 0
PREFIX = "first_stage_model."


def _put(store, name, t, whole=False):
    if whole:
        store[name] = t.detach().float().numpy()
        return
    s = synth.summarize(t)
    store[name + "/sample"] = s["sample"].numpy()
    store[name + "/shape"] = np.asarray(s["shape"], dtype=np.int64)
    store[name + "/moments"] = np.asarray([s["mean"], s["std"], s["l2"], s["absmax"]], dtype=np.float64)


vae_inputs = vae_oracle.vae_inputs


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    ref_shim.load_reference()
    ae_mod = importlib.import_module("model_lib.ControlNet.ldm.models.autoencoder")
    dm_mod = importlib.import_module("model_lib.ControlNet.ldm.modules.diffusionmodules.model")
    # diffusionmodules/model.py imports xformers unconditionally and then prefers MemoryEfficientAttnBlock
    # (model.py:282-283); xformers is a stub here, so select the vanilla AttnBlock (model.py:152-203): same
    # parameters, same mathematics softmax(q k^T c^-0.5) v — a run-time flag of the module, no source change
    dm_mod.XFORMERS_IS_AVAILBLE = False
    dist_mod = importlib.import_module("model_lib.ControlNet.ldm.modules.distributions.distributions")
    cfg = ref_shim.load_yaml()
    fs = cfg["model"]["params"]["first_stage_config"]["params"]
    vae = ae_mod.AutoencoderKL(ddconfig=dict(fs["ddconfig"]), lossconfig=fs["lossconfig"], embed_dim=fs["embed_dim"])
    vae.eval()
    manifest = {PREFIX + k: list(v.shape) for k, v in vae.state_dict().items()}
    with open(MANIFEST, "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    print(f"{len(manifest)} VAE tensors, {sum(int(np.prod(s)) for s in manifest.values()) / 1e6:.1f} M parameters")
    sd = synth.synth_state_dict(manifest, SEED)
    vae.load_state_dict({k[len(PREFIX):]: v for k, v in sd.items()}, strict=True)
    scale = float(cfg["model"]["params"]["scale_factor"])
    assert abs(scale - vae_oracle.SCALE_FACTOR) < 1e-12

    def ref_decode_first_stage(z):  # ddpm.py:2107-2108
        return vae.decode(1.0 / scale * z)

    # ---- small: B=2, latent 16 -> image 128 ----
    store = {}
    z, img, noise = vae_inputs(2, 16)
    taps = {}
    hooks = [vae.decoder.mid.block_2.register_forward_hook(lambda m, i, o: taps.__setitem__("dec/mid", o))]
    for lvl in range(4):
        last = vae.decoder.up[lvl].upsample if lvl != 0 else vae.decoder.up[lvl].block[2]
        hooks.append(last.register_forward_hook(lambda m, i, o, lvl=lvl: taps.__setitem__(f"dec/up{lvl}", o)))
    for lvl in range(4):
        last = vae.encoder.down[lvl].downsample if lvl != 3 else vae.encoder.down[lvl].block[1]
        hooks.append(last.register_forward_hook(lambda m, i, o, lvl=lvl: taps.__setitem__(f"enc/down{lvl}", o)))
    t0 = time.time()
    dec = ref_decode_first_stage(z)
    post = vae.encode(img)
    assert isinstance(post, dist_mod.DiagonalGaussianDistribution)
    print(f"small decode+encode: {time.time() - t0:.1f}s; image std {float(dec.std()):.3f}, moments std {float(post.parameters.std()):.3f}")
    for h in hooks:
        h.remove()
    _put(store, "vae16/decoded", dec, whole=True)
    _put(store, "vae16/moments", post.parameters, whole=True)
    # the posterior sample with a fixed noise tensor (distributions.py:35-37 draws its own; same formula)
    _put(store, "vae16/encoding", scale * (post.mean + post.std * noise), whole=True)
    for k, v in taps.items():
        _put(store, "vae16/" + k, v)
    np.savez_compressed(os.path.join(GOLDEN, "vae16.npz"), **store)

    # ---- headline size: B=1, latent 64 -> image 512 ----
    store = {}
    z, _, _ = vae_inputs(1, 64)
    t0 = time.time()
    dec = ref_decode_first_stage(z)
    sec = time.time() - t0
    print(f"full decode 64x64 -> 512x512: {sec:.1f}s on {torch.get_num_threads()} threads")
    _put(store, "vae64/decoded", dec)
    store["vae64/decode_seconds"] = np.asarray([sec, torch.get_num_threads()], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "vae64.npz"), **store)
    print("wrote", MANIFEST, "and tests/golden/vae16.npz, vae64.npz")


if __name__ == "__main__":
    main()
