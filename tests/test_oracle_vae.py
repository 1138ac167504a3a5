"""Pins the CPU restatement of the first-stage autoencoder (oracle/vae_restatement.py, SURVEY §8f rank 2 — the
row after the hot path) against golden vectors produced by the UNMODIFIED reference AutoencoderKL
(oracle/make_golden_vae.py, run in the build container).  CPU only, fp32."""
import json
import os

import pytest
import torch

from oracle import synth
from oracle import vae_restatement as V
from tests import golden_util as G

TOL = 2e-4  # fp32 vs fp32, different op ordering only
MANIFEST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "magicdance_b200", "vae_manifest.json")


@pytest.fixture(scope="module")
def weights():
    torch.set_grad_enabled(False)
    with open(MANIFEST) as f:
        manifest = json.load(f)
    assert len(manifest) == 248  # first_stage_model.* tensors of the released checkpoint
    assert all(k.startswith(V.PREFIX) for k in manifest)
    return synth.synth_state_dict(manifest, seed=0)


def test_decode_first_stage_small_matches_reference(weights):
    g = G.load("vae16")
    z, _, _ = V.vae_inputs(2, 16)
    taps = {}
    img = V.decode_first_stage(weights, z, taps)
    assert tuple(img.shape) == (2, 3, 128, 128)
    assert G.rel_l2(img, torch.from_numpy(g["vae16/decoded"])) <= TOL
    G.check_summary(g, "vae16/dec/mid", taps["mid"], TOL)
    for lvl in range(4):
        G.check_summary(g, f"vae16/dec/up{lvl}", taps[f"up{lvl}"], TOL)
    assert float(img.std()) > 0.1  # not vacuous


def test_encode_first_stage_small_matches_reference(weights):
    g = G.load("vae16")
    _, img, noise = V.vae_inputs(2, 16)
    taps = {}
    moments = V.vae_encode_moments(weights, img, taps)
    assert tuple(moments.shape) == (2, 8, 16, 16)
    assert G.rel_l2(moments, torch.from_numpy(g["vae16/moments"])) <= TOL
    for lvl in range(4):
        G.check_summary(g, f"vae16/enc/down{lvl}", taps[f"down{lvl}"], TOL)
    enc = V.get_first_stage_encoding(moments, noise)
    assert G.rel_l2(enc, torch.from_numpy(g["vae16/encoding"])) <= TOL
    # the posterior mode is the mean half, scaled
    assert torch.equal(V.get_first_stage_encoding(moments), V.SCALE_FACTOR * moments[:, :4])


def test_decode_first_stage_headline_size_matches_reference(weights):
    g = G.load("vae64")
    z, _, _ = V.vae_inputs(1, 64)
    img = V.decode_first_stage(weights, z)
    assert tuple(img.shape) == (1, 3, 512, 512)
    G.check_summary(g, "vae64/decoded", img, TOL)
