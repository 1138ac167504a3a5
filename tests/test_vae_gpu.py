"""First-stage VAE (the row right after the hot path) on the GPU: magicdance_b200/vae.py and the
drop-in AutoencoderKL against the goldens of the UNMODIFIED reference AutoencoderKL (oracle/make_golden_vae.py;
ldm/models/autoencoder.py:83-91, ldm/modules/diffusionmodules/model.py:452-655).  fp16 storage / fp32 accumulation
vs fp32: rel-L2 <= 5e-3."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 5e-3


@pytest.fixture(scope="module")
def vae_sd():
    from magicdance_b200 import synth
    torch.set_grad_enabled(False)
    with open(os.path.join(REPO, "magicdance_b200", "vae_manifest.json")) as f:
        manifest = json.load(f)
    return synth.synth_state_dict(manifest, seed=0)


def test_decoder_matches_reference_goldens(vae_sd):
    from magicdance_b200.vae import PackedVaeDecoder, VaeDecoder
    from oracle import vae_restatement as V  # inputs only
    from tests import golden_util as G
    dec = VaeDecoder(PackedVaeDecoder(vae_sd, "cuda"))
    g = G.load("vae16")
    z, _, _ = V.vae_inputs(2, 16)
    img = dec.decode(z.cuda())
    assert tuple(img.shape) == (2, 3, 128, 128) and img.dtype == torch.float32
    assert G.rel_l2(img, torch.from_numpy(g["vae16/decoded"])) <= TOL
    g = G.load("vae64")  # the headline size: 64x64 latent -> 512x512 image (stored as subsample + moments)
    z, _, _ = V.vae_inputs(1, 64)
    img = dec.decode(z.cuda())
    assert tuple(img.shape) == (1, 3, 512, 512)
    G.check_summary(g, "vae64/decoded", img, TOL)


def test_encoder_matches_reference_goldens(vae_sd):
    from magicdance_b200.vae import PackedVaeEncoder, VaeEncoder
    from oracle import vae_restatement as V
    from tests import golden_util as G
    enc = VaeEncoder(PackedVaeEncoder(vae_sd, "cuda"))
    g = G.load("vae16")
    _, img, noise = V.vae_inputs(2, 16)
    mom = enc.encode(img.cuda())
    assert tuple(mom.shape) == (2, 8, 16, 16)
    assert G.rel_l2(mom, torch.from_numpy(g["vae16/moments"])) <= TOL


def test_dropin_autoencoder_behind_the_ldm_api(vae_sd):
    """create_model(yaml) resolves first_stage_config to the drop-in AutoencoderKL; decode_first_stage /
    encode_first_stage / get_first_stage_encoding (test_tiktok.py:269-272) give the reference's results."""
    from magicdance_b200 import synth
    from magicdance_b200.dropin.autoencoder import AutoencoderKL
    from model_lib.ControlNet.cldm.model import create_model
    from oracle import vae_restatement as V
    from tests import golden_util as G
    m = create_model(os.path.join(REPO, "model_lib", "ControlNet", "models", "cldm_v15_reference_only_pose.yaml"))
    assert isinstance(m.first_stage_model, AutoencoderKL)
    sd = dict(synth.synth_state_dict(seed=0))
    sd.update(vae_sd)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= set(synth.SCHEDULE_KEYS)
    m = m.cuda().eval()
    g = G.load("vae16")
    z, img, noise = V.vae_inputs(2, 16)
    out = m.decode_first_stage(z.cuda())
    assert G.rel_l2(out, torch.from_numpy(g["vae16/decoded"])) <= TOL
    post = m.encode_first_stage(img.cuda())
    assert G.rel_l2(post.parameters, torch.from_numpy(g["vae16/moments"])) <= TOL
    enc = m.get_first_stage_encoding(post.mode())
    assert G.rel_l2(enc, V.SCALE_FACTOR * torch.from_numpy(g["vae16/moments"][:, :4])) <= TOL
